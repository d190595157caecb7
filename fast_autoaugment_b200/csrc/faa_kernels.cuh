// faa_kernels.cuh - launch-side declarations shared by faa_kernels.cu and faa_cabi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "faa_core.cuh"

namespace faa {

enum OutType : int32_t { OUT_F16 = 0, OUT_BF16 = 1, OUT_F32 = 2, OUT_U8_HWC = 3 };

// step 1 (one thread per image): decisions -> per-image program
struct ResolveParams {
    const OpRec* ops;           // compiled policy [n_sub][n_op][2]
    const double* probs;        // [n_sub][n_op]
    const Sample* samples;      // [n] resolved decisions, or nullptr -> draw with Philox
    const Box* boxes;           // [n][n_op] (may be nullptr when no Cutout box is needed)
    Prog* progs;                // [n] out
    int32_t* order;             // [n] out: local image indices, most expensive first (optional)
    int32_t* n_heavy;           // out (optional): [0] = end of the heavy segment of the order, [1] = end of the mid segment
    int32_t split;              // 1: sort light programs behind heavy ones (two pixel launches); 2: heavy | mid | light
    Sample* samples_out;        // optional [n]
    Box* boxes_out;             // optional [n][n_op]
    RngCfg rng;
    int32_t first, n;           // images [first, first+n) of the arrays above
    int32_t H, W, out_h, out_w, n_sub, n_op, op_base, apply_tail;
    int32_t allow;              // bit 0: the pixel kernel has a materialisation chunk, bit 1: a global scratch image
    int32_t* ready;             // optional: set to `ticket` (release) once programs / order / n_heavy are written
    int32_t ticket;
    int32_t pdl;                // launch with programmatic stream serialization (chained steps)
    const uint32_t* wait_done;  // optional: before writing anything, spin until *wait_done has reached wait_target (the pixel
    uint32_t wait_target;       // CTAs that read this slot's previous programs count themselves there when they finish)
};
cudaError_t launch_resolve(const ResolveParams& p, cudaStream_t stream);

// row-band geometry of a pixel launch, precomputed on the host (no divisions in the kernels)
struct BandGeom {
    int32_t bands;              // CTAs per image
    int32_t band_cap;           // bytes of dynamic shared memory per staged band (0: staging off)
    int32_t y[9];               // band b owns image rows [y[b], y[b+1]) for whole-image statistics
    int32_t oy[9];              // ... and output rows [oy[b], oy[b+1])
    uint32_t lo[8], len[8];     // staged byte range of the raw image per band (len 0: nothing staged)
};

// step 2 (one cluster per image): pixels
struct AugParams {
    const uint8_t* in;          // [n_all][H][W][3] uint8
    void* out;                  // [B][3][out_h][out_w] OutT  (or [B][out_h][out_w][3] uint8)
    const Prog* progs;          // [n_all]
    const int32_t* partner;     // [B] index into [0, n_all) or nullptr (no mixup)
    const int32_t* order;       // [n_all] LPT schedule written by the resolve kernel, or nullptr
    const int32_t* n_heavy;     // device counters: schedule entries [0, n_heavy[0]) -> cluster kernel, [n_heavy[0], n_heavy[1]) -> mid
                                // kernel (empty in a two-way split), rest -> light kernel; nullptr = no split
    const float* norm_tab;      // [3][256] exact fp32 ToTensor+Normalize values ([n_all][3][256] when norm_stride != 0)
    int32_t norm_stride;        // 768: one table per image (Lighting, augmentations.py:197-215); 0: one table per launch
    uint8_t* scratch;           // [n_all][H][W][3] uint8 scratch image for Sharpness->gather programs, or nullptr
    int32_t B, H, W, out_h, out_w;
    int32_t first;              // index of this launch's image 0 inside the n_all arrays
    int32_t in_mod;             // > 0: replicated launch (TTA): entry v reads input image v % in_mod
    int32_t use_zero_box;
    int32_t bands;              // CTAs (== cluster size) per image of the cluster kernel (== geo[0].bands)
    BandGeom geo[2];            // [0] cluster kernel, [1] light streaming kernel
    uint32_t rcp_out_qpr, rcp_w, rcp_wq;   // 2^32 / ceil(out_w/4), 2^32 / W, 2^32 / (W/4)  (rounded up) for fastdiv
    uint32_t rcp_opr;           // 2^32 / (W/8) (octet fast paths; 0 when W % 8 != 0)
    int32_t octets;             // 1: 8-pixel fast paths allowed (W % 8 == 0, out size == image size, float output 16-byte aligned)
    int32_t stage;              // 1: TMA-stage the raw row band into shared memory
    int32_t band_cap;           // bytes of dynamic shared memory per staged band
    int32_t crop_pad;           // max |crop_dy| (RandomCrop padding)
    int32_t mat_cap;            // bytes of the materialisation chunk (0: none), a whole number of rows >= 3
    int32_t pdl;                // launched with programmatic stream serialization
    const int32_t* ready;       // chained steps: spin until *ready == ticket before reading programs / order / n_heavy
    int32_t ticket;
    int32_t chain;              // 1: steps are chained with programmatic dependent launches on ONE stream - no
                                // griddepcontrol.wait (nothing of the previous kernel is consumed), trigger the
                                // dependents once this CTA has copied its program
                                // 2: same, with persistent mid / light kernels - the dependents are released at once and
                                // the slot's next writer waits for `done` instead
    int32_t grid_y;             // > 0: launch this many schedule rows (CTAs / clusters per band); each one loops over the
                                // entries row, row + grid_y, ... of its segment.  0: one row per image (B)
    uint32_t* done;             // optional completion counter: every CTA adds 1 when it has finished (release)
    // self-resolving launches (small batches, cluster kernel only): no resolve kernel, no program array - thread 0 of each
    // CTA draws its image's decisions (Philox, the same counters as faa_resolve_kernel) and builds the program itself
    int32_t self_resolve;
    const OpRec* sr_ops; const double* sr_probs;
    RngCfg sr_rng;
    int32_t sr_n_sub, sr_n_op, sr_op_base, sr_apply_tail, sr_allow;
    float scale[3], bias[3];
    float lam, one_minus_lam;   // mixup weights (fp32 of the Python floats)
};
// number of CTAs launch_augment starts for (p, which): what `done` advances by
unsigned augment_cta_count(const AugParams& p, int which);
int resident_ctas_per_sm(int which);     // the launch bounds of the light (1) / mid (2) / cluster (0) kernel
cudaError_t launch_augment(const AugParams& p, int out_type, bool use_tab, int which, cudaStream_t stream);   // which: 0 cluster, 1 light, 2 mid

cudaError_t launch_mixup(const void* data, void* out, const int64_t* perm, int batch, int64_t n_per_sample,
                         int dtype, float lam, float one_minus_lam, cudaStream_t stream);

cudaError_t launch_mix_u8(const uint8_t* a, const uint8_t* b, const int32_t* partner, const int16_t* zb_a, const int16_t* zb_b,
                          const float* norm_tab, void* out, int batch, int H, int W, int dtype, float lam, float one_minus_lam,
                          cudaStream_t stream, const uint8_t* const* b_ptrs = nullptr);

cudaError_t launch_color_jitter(const uint8_t* in, uint8_t* out, const void* recs, int batch, int H, int W, cudaStream_t stream);
cudaError_t launch_lighting_tables(const float* rgb, float* tabs, int n, const float mean[3], const float std[3], cudaStream_t stream);

int pick_bands(int H, int W, int out_h, int out_w);
void fill_geom(BandGeom& g, int bands, int H, int W, int out_h, int crop_pad, bool stage);
uint32_t band_capacity(int bands, int H, int W, int out_h, int crop_pad);

}  // namespace faa
