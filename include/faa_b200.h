/*
 * faa_b200.h - C ABI of the B200-native Fast AutoAugment augmentation hot path.
 *
 * Plain C: opaque handle, plain pointers and sizes, int status returns, no C++
 * or torch types, no exceptions across the boundary.  Every device entry point
 * takes an explicit CUDA stream (passed as void*, i.e. a cudaStream_t) and is
 * asynchronous with respect to the host; the caller owns every device buffer,
 * the library owns only the policy handle and its small device-side tables.
 *
 * The reference (kakaobrain/fast-autoaugment @ 2424224) has no FFI: its
 * boundary for this path is a set of Python callables.  Each entry point below
 * names the reference interface it replaces (file:line relative to the
 * reference root).  The Python mirror of that surface lives in
 * fast_autoaugment_b200/ and binds this header with ctypes; INTEGRATION.md
 * shows the binding a reference maintainer would add.
 */
#ifndef FAA_B200_H
#define FAA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAA_ABI_VERSION 1
#define FAA_MAX_FUSED_OPS 2      /* ops of one sub-policy applied by one launch */
#define FAA_MAX_POLICY_OPS 8     /* ops per sub-policy (search.py --num-op); >2 runs as chained launches */
#define FAA_MAX_DIM 8192         /* max H or W */

/* status codes; the Python layer maps them to the reference's exceptions */
enum faa_status {
    FAA_OK = 0,
    FAA_ERR_UNKNOWN_OP = 1,   /* KeyError      - augmentations.py:189 get_augment */
    FAA_ERR_MAGNITUDE = 2,    /* AssertionError- augmentations.py:14,21,28,36,44,51,58,81,86,92,98,103,108,113,118 */
    FAA_ERR_VALUE = 3,        /* ValueError    - bad argument */
    FAA_ERR_CUDA = 4,         /* RuntimeError  - CUDA runtime error, see faa_last_error() */
    FAA_ERR_NO_DEVICE = 5,    /* RuntimeError  - no CUDA device: there is NO CPU fallback */
    FAA_ERR_UNSUPPORTED = 6   /* RuntimeError  - shape / option outside what the kernels handle */
};

/* op ids = position in augment_list(for_autoaug=True), augmentations.py:156-182 */
enum faa_op_id {
    FAA_SHEAR_X = 0, FAA_SHEAR_Y = 1, FAA_TRANSLATE_X = 2, FAA_TRANSLATE_Y = 3, FAA_ROTATE = 4,
    FAA_AUTOCONTRAST = 5, FAA_INVERT = 6, FAA_EQUALIZE = 7, FAA_SOLARIZE = 8, FAA_POSTERIZE = 9,
    FAA_CONTRAST = 10, FAA_COLOR = 11, FAA_BRIGHTNESS = 12, FAA_SHARPNESS = 13, FAA_CUTOUT = 14,
    FAA_CUTOUT_ABS = 15, FAA_POSTERIZE2 = 16, FAA_TRANSLATE_X_ABS = 17, FAA_TRANSLATE_Y_ABS = 18,
    FAA_NUM_OPS = 19
};

/* which random draws an applied op consumes (augmentations.py:15,22,29,37,45,52,59,131-132) */
enum faa_draw { FAA_DRAW_NONE = 0, FAA_DRAW_MIRROR = 1, FAA_DRAW_BOX = 2 };

enum faa_dtype { FAA_F16 = 0, FAA_BF16 = 1, FAA_F32 = 2, FAA_U8_HWC = 3 };

/*
 * Per-sample resolved decisions: what the reference's RNG draws decided for one
 * image (data.py:257-264 Augmentation.__call__, torchvision RandomCrop /
 * RandomHorizontalFlip data.py:40-41, CutoutDefault data.py:235-244).
 * Filled on the host by the parity sampler or on the device by the Philox one.
 */
typedef struct faa_sample {
    uint16_t sub;          /* chosen sub-policy (random.choice, data.py:259)              */
    uint8_t  gate;         /* bit j: op j passed its `random.random() > pr` gate (:261)    */
    uint8_t  sign;         /* bit j: op j drew the mirrored (-v) variant                   */
    int8_t   crop_dy;      /* RandomCrop: top  - padding  (source row = out row + crop_dy) */
    int8_t   crop_dx;      /* RandomCrop: left - padding                                   */
    uint8_t  flip;         /* RandomHorizontalFlip fired                                   */
    uint8_t  reserved;
    int16_t  zero_box[4];  /* CutoutDefault half-open, clipped: y1, y2, x1, x2 (data.py:241-246) */
} faa_sample_t;            /* 16 bytes */

/* per-sample, per-op inclusive Cutout rectangle x0, y0, x1, y1 (augmentations.py:134-143),
 * already truncated to integers, NOT yet clipped to the image */
typedef struct faa_box { int16_t x0, y0, x1, y1; } faa_box_t;

/* the part of the train chain behind the policy (data.py:39-44, 64, 70-72, 111-112) */
typedef struct faa_tail {
    int32_t out_h, out_w;     /* RandomCrop size (== H, W when there is no crop)            */
    int32_t out_dtype;        /* enum faa_dtype; FAA_U8_HWC skips ToTensor/Normalize         */
    int32_t use_zero_box;     /* apply faa_sample.zero_box (CutoutDefault)                   */
    float   mean[3], std[3];  /* Normalize                                                  */
    int32_t crop_pad;         /* RandomCrop padding = bound on |crop_dy| (sizes the staged band; a hint) */
    int32_t reserved;
} faa_tail_t;

/* parameters of the device-side (Philox4x32-10) sampler: the distribution of every draw
 * is the reference's, the stream is not (statistical equivalence, not replay) */
typedef struct faa_rng {
    uint64_t seed;            /* key                                                        */
    uint64_t first_index;     /* global index of sample 0 of this call (shard offset)        */
    int32_t  crop_pad;        /* RandomCrop padding (0 = no crop)                            */
    int32_t  hflip;           /* RandomHorizontalFlip present                                */
    int32_t  zero_box_len;    /* CutoutDefault length (0 = off)                              */
    int32_t  reserved;
} faa_rng_t;

typedef struct faa_policy faa_policy_t;

/* ---- misc ------------------------------------------------------------------ */
int         faa_abi_version(void);
const char* faa_last_error(void);          /* thread-local message of the last failing call */
int         faa_device_count(void);        /* 0 when no usable CUDA device */
int         faa_op_id_from_name(const char* name);   /* -1 if unknown; names of augment_list() */
const char* faa_op_name(int op_id);
int         faa_op_range(int op_id, double* low, double* high);   /* augmentations.py:157-181 */

/* ---- policy: replaces Augmentation.__init__ (data.py:254-255) over the archive.py
 *      list-of-sub-policies format.  ops/probs/levels are row-major [n_sub][n_op].
 *      Like the reference, an unknown op / out-of-range magnitude is only an error when the op
 *      is actually applied: the host sampler reports it then (KeyError / AssertionError); the
 *      device sampler, which cannot raise, refuses such a policy up front.            */
int faa_policy_create(const int32_t* op_ids, const double* probs, const double* levels,
                      int n_sub, int n_op, faa_policy_t** out);
int faa_policy_destroy(faa_policy_t* p);
int faa_policy_dims(const faa_policy_t* p, int* n_sub, int* n_op);
/* host-side compiled op record (32 bytes) for (sub, op, sign) at image size (h, w):
 * exposes the level->magnitude->fixed-point compilation (augmentations.py:192-194 + Pillow
 * matrix set-up) for tests and foreign hosts.  No GPU needed. */
int faa_policy_compiled_op(faa_policy_t* p, int h, int w, int sub, int op, int sign, int32_t out8[8]);
int faa_policy_draw_kind(const faa_policy_t* p, int sub, int op);     /* enum faa_draw */
/* CutoutAbs box from the two uniforms (augmentations.py:131-137), host helper */
int faa_cutout_box(const faa_policy_t* p, int h, int w, int sub, int op, double ux, double uy,
                   faa_box_t* out);

/* ---- host parity sampler: replays Augmentation.__call__'s draws (data.py:257-264) for
 *      `batch` images drawn one after another, from explicit MT19937 states of Python's
 *      `random` (624 words + index) and numpy's legacy global RandomState (same layout).
 *      States are advanced in place.  Tail draws (torch generator) are not covered here. */
int faa_sample_policy_mt(const faa_policy_t* p, int batch, int h, int w,
                         uint32_t py_state[625], uint32_t np_state[625],
                         faa_sample_t* out_samples, faa_box_t* out_boxes /* [batch][n_op] */);

/* ---- device sampler (Philox): fills samples/boxes on the device -------------- */
int faa_sample_philox(faa_policy_t* p, int batch, int h, int w, const faa_tail_t* tail,
                      const faa_rng_t* rng, faa_sample_t* d_samples, faa_box_t* d_boxes,
                      void* stream);

/* ---- the hot path: replaces, for a whole batch, Augmentation.__call__ (data.py:257-264)
 *      -> apply_augment (augmentations.py:192-194) -> the 19 ops (augmentations.py:13-144)
 *      -> RandomCrop/HFlip/ToTensor/Normalize (data.py:40-43, 64, 70-72)
 *      -> CutoutDefault (data.py:235-250).
 *      d_in : uint8 [batch][h][w][3] (HWC, contiguous) on the device
 *      d_out: [batch][3][out_h][out_w] of tail->out_dtype (or uint8 HWC)
 *      d_samples / d_boxes: resolved decisions; if d_samples == NULL the kernel draws them
 *      itself from `rng` (fused Philox mode).  op_base selects which FAA_MAX_FUSED_OPS-wide
 *      window of the sub-policy this launch applies (chained launches for n_op > 2). */
int faa_augment(faa_policy_t* p, const uint8_t* d_in, void* d_out, int batch, int h, int w,
                const faa_tail_t* tail, const faa_sample_t* d_samples, const faa_box_t* d_boxes,
                const faa_rng_t* rng, int op_base, void* stream);

/* fused Mixup variant: out[i] = lam*aug(in[i]) + one_minus_lam*aug(in[partner[i]]) in fp32
 * (lam and one_minus_lam are the fp32 casts of the Python floats lam and 1-lam)
 * (aug_mixup.py:13-23 with the pairing resolved by the caller; partner indexes d_in_all,
 * which may be an all-gathered or peer-mapped array of n_all images with its own
 * samples/boxes).  d_partner == NULL or lam == 1 degenerates to faa_augment. */
int faa_augment_mixup(faa_policy_t* p, const uint8_t* d_in_all, int n_all, int first, void* d_out,
                      int batch, int h, int w, const faa_tail_t* tail,
                      const faa_sample_t* d_samples_all, const faa_box_t* d_boxes_all,
                      const faa_rng_t* rng, const int32_t* d_partner, float lam, float one_minus_lam,
                      void* stream);

/* ---- overlap of consecutive calls.  The kernels of one call are chained with programmatic dependent launches and
 * overlap each other; with on != 0 the kernels of call N+1 may also start while call N (same handle, same stream,
 * disjoint buffers) is still running - worth ~15 % at 224x224 b512.  The caller promises that the INPUT batch of every
 * call was complete before the previous call on that stream was issued (e.g. device-resident data, or a producer that
 * runs one batch ahead): the first kernel of an overlapping call does not wait for the kernel right in front of it in
 * the stream.  Default: off (the first kernel of every call is an ordinary stream-ordered launch). */
int faa_policy_set_overlap(faa_policy_t* p, int on);

/* ---- several consecutive batches in one call: the loop `for data, label in loader:` of train.py:47-49 when the
 * dataset is device-resident (data.py:114-224 replaced by a DeviceDataset) - the caller knows the next n_steps
 * batches in advance, so their launches are issued back to back without returning to the interpreter (small-image
 * steps are bound by the host's launch rate otherwise).  Step k reads d_in[k] ([batch][h][w][3] uint8), writes
 * d_out[k] and draws the decisions of global samples rng->first_index + k*index_stride + i; it equals
 * faa_augment(..., rng with that first_index, ...).  Policies of at most FAA_MAX_FUSED_OPS ops. */
int faa_augment_many(faa_policy_t* p, int n_steps, const uint8_t* const* d_in, void* const* d_out,
                     int batch, int h, int w, const faa_tail_t* tail, const faa_rng_t* rng,
                     uint64_t index_stride, void* stream);

/* ---- test-time-augmentation batching: replaces the `num_policy` validation loaders of
 * eval_tta (search.py:87-90: one get_dataloaders call per replica, each drawing its own
 * sub-policies for the SAME validation batch, the losses reduced per sample at :116-125).
 * ONE launch augments d_in [batch] `replicas` times: d_out [replicas][batch][3][out_h][out_w];
 * replica r uses the decisions of global samples rng->first_index + r*batch + i, i.e. it equals
 * faa_augment called with first_index + r*batch.  Policies of at most FAA_MAX_FUSED_OPS ops. */
int faa_augment_tta(faa_policy_t* p, const uint8_t* d_in, void* d_out, int batch, int replicas,
                    int h, int w, const faa_tail_t* tail, const faa_rng_t* rng, void* stream);

/* same call with HOST buffers: pinned staging, chunked H2D / kernel / D2H pipeline inside.
 * h_out may be NULL (result stays on the device in d_out_keep, which may also be NULL). */
int faa_augment_host(faa_policy_t* p, const uint8_t* h_in, void* h_out, void* d_out_keep,
                     int batch, int h, int w, const faa_tail_t* tail, const faa_rng_t* rng,
                     void* stream);

/* ---- standalone Mixup on already-augmented device tensors: replaces mixup()
 *      (aug_mixup.py:13-23) given the resolved permutation and lambda.
 *      out[i] = data[i]*lam + data[perm[i]]*(1-lam), fp32 math, n_per_sample elements each. */
int faa_mixup(const void* d_data, void* d_out, const int64_t* d_perm, int batch,
              int64_t n_per_sample, int dtype, float lam, float one_minus_lam, void* stream);

/* ---- Mixup of AUGMENTED uint8 images (the multi-GPU / large-batch route): the policy part of the
 * chain ran with a uint8 HWC output (faa_augment with tail.out_dtype = FAA_U8_HWC: policy + crop +
 * flip), possibly on another GPU; this call finishes data.py:42-43 (ToTensor, Normalize), data.py:228-250
 * (CutoutDefault: one half-open zero box [y0,y1)x[x0,x1) per SOURCE image, int16[4] each, may be NULL) and
 * aug_mixup.py:21 in one streaming pass:
 *     d_out[i] = norm(d_a[i]) * lam + norm(d_b[d_partner[i]]) * one_minus_lam      (fp32, then tail->out_dtype)
 * Same values as faa_augment_mixup on the raw images.  W % 4 == 0. */
int faa_mix_u8(faa_policy_t* p, const uint8_t* d_a, const uint8_t* d_b, const int32_t* d_partner,
               const int16_t* d_zero_box_a, const int16_t* d_zero_box_b, void* d_out, int batch, int h, int w,
               const faa_tail_t* tail, float lam, float one_minus_lam, void* stream);

/* ---- the same pass with the partner exchange INSIDE the kernel: d_partner_ptrs[i] (device array of `batch` pointers)
 * is the address of sample i's partner image - local, or in the memory of another GPU of the node mapped into this
 * process (CUDA IPC / symmetric memory; NVLink peer access enabled).  The partner's bytes travel over NVSwitch as the
 * kernel's loads: no all-to-all, no staging copy.  The caller orders it behind the partners' augmentation (a barrier
 * across the ranks) and keeps their buffers alive until it has run.  d_zero_box_b: one box PER SAMPLE (its partner's). */
int faa_enable_peer_access(int peer_device);      /* cudaDeviceEnablePeerAccess from the current device (idempotent) */
/* a device buffer that the other processes of the node can map (cudaMalloc + cudaIpcGetMemHandle; 64-byte handle),
 * the mapping of such a buffer under the current device (cudaIpcOpenMemHandle, lazy peer access), and their release */
int faa_peer_alloc(size_t bytes, void** d_ptr, unsigned char* handle64);
int faa_peer_open(const unsigned char* handle64, void** d_ptr);
int faa_peer_close(void* d_ptr);
int faa_peer_free(void* d_ptr);
int faa_mix_u8_peer(faa_policy_t* p, const uint8_t* d_a, const uint8_t* const* d_partner_ptrs,
                    const int16_t* d_zero_box_a, const int16_t* d_zero_box_b, void* d_out, int batch, int h, int w,
                    const faa_tail_t* tail, float lam, float one_minus_lam, void* stream);

/* ---- ImageNet train chain pieces (data.py:60-73), "next" row N2 --------------------------------
 * torchvision ColorJitter(brightness, contrast, saturation) (data.py:65-69) on uint8 HWC images, in place
 * allowed (d_out == d_in): per image the ops of order[] (torch.randperm(4): 0 brightness, 1 contrast,
 * 2 saturation, 3 hue = absent) with the factors alpha[] as PIL ImageEnhance blends - the arithmetic of
 * the policy ops Brightness / Contrast / Color with per-image magnitudes. */
typedef struct faa_jitter { float alpha[3]; uint8_t order[4]; } faa_jitter_t;
int faa_color_jitter(const uint8_t* d_in, uint8_t* d_out, int batch, int h, int w,
                     const faa_jitter_t* d_recs, void* stream);

/* Lighting (augmentations.py:197-215) sits between ToTensor and Normalize (data.py:70-72):
 * d_rgb [n][3] fp32 (device) = the per-image offsets eigvec . (alpha * eigval); subsequent faa_augment
 * launches over n images normalise with per-image tables ((u8/255 + rgb[c]) - mean[c]) / std[c] in torch's
 * fp32 operation order.  NULL switches it off again.  The array must stay valid until those launches ran. */
int faa_policy_set_lighting(faa_policy_t* p, const float* d_rgb, int n);

/* number of kernels this library has launched since load (bench bookkeeping) */
uint64_t faa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FAA_B200_H */
