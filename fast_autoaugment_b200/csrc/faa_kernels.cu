// faa_kernels.cu - sm_100a kernels of the augmentation hot path.
//
// One thread-block CLUSTER per image: the image's rows are split into `bands` row bands,
// one CTA each (cluster dims = (bands,1,1), grid = (bands, batch)).  Every CTA
//   1. resolves the image's decisions (sub-policy, gates, signs, boxes, crop/flip/zero-box)
//      from the per-sample record - or draws them itself (fused Philox mode),
//   2. for each op that needs whole-image statistics (AutoContrast / Equalize histogram,
//      Contrast mean luma) scans its band of the *intermediate* image, reduces the partial
//      statistics across the cluster through distributed shared memory, and builds the
//      3x256-byte LUT of that op,
//   3. streams its band of the OUTPUT: tail index map (zero box, flip, crop) -> lazy
//      evaluation of the op chain back to the raw uint8 pixels -> ToTensor+Normalize ->
//      NCHW fp16/bf16/fp32 (or uint8 HWC) vector stores.
// With a Mixup partner the same evaluation runs for the partner image and the two
// normalised values are mixed in fp32 before the store (aug_mixup.py:21).
//
// Algorithmic HBM bytes per image: 3*H*W read + out_elem*3*out_h*out_w written.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "faa_kernels.cuh"

namespace cg = cooperative_groups;

namespace faa {

constexpr int kThreads = 256;

struct __align__(16) ImgState {
    uint32_t hist[2][768];      // per-slot local partial histograms (read remotely through DSMEM)
    uint32_t tot[768];          // cluster-reduced histogram of the slot being built
    uint8_t lut[2][768];
    HistPart parts[3][32];
    unsigned long long suml[2]; // per-slot local partial luma sums
    Sample smp;
    Box box[2];
    OpRec op[2];
};

struct FastDiv {
    uint32_t d, rcp;
    __device__ __forceinline__ void init(uint32_t dd) { d = dd; rcp = recip32(dd); }
    __device__ __forceinline__ uint32_t div(uint32_t q) const { return d == 1u ? q : fastdiv(q, rcp); }
};

// ---------------------------------------------------------------------------------------
// step 1: decisions -> smem
__device__ void load_program(const AugParams& P, int idx, ImgState& st) {
    if (threadIdx.x == 0) {
        Box bx[8];
        Sample s;
        if (P.samples != nullptr) {
            s = P.samples[idx];
            for (int j = 0; j < 2; ++j) {
                int jj = P.op_base + j;
                if (jj < P.n_op && P.boxes != nullptr) bx[jj] = P.boxes[(size_t)idx * P.n_op + jj];
                else if (jj < 8) { bx[jj].x0 = bx[jj].y0 = 0; bx[jj].x1 = bx[jj].y1 = -1; }
            }
        } else {
            philox_sample(P.rng, P.rng.first_index + (uint64_t)idx, P.ops, P.probs, P.n_sub, P.n_op,
                          P.H, P.W, P.out_h, P.out_w, s, bx);
        }
        if (!P.apply_tail) { s.crop_dx = s.crop_dy = 0; s.flip = 0; }
        st.smp = s;
        for (int j = 0; j < 2; ++j) {
            int jj = P.op_base + j;
            OpRec o; o.kind = K_NONE; o.a[0] = o.a[1] = o.a[2] = o.a[3] = o.a[4] = o.a[5] = 0; o.draw = 0;
            if (jj < P.n_op && ((s.gate >> jj) & 1u))
                o = P.ops[((size_t)s.sub * P.n_op + jj) * 2 + ((s.sign >> jj) & 1u)];
            st.op[j] = o;
            Box b; b.x0 = b.y0 = 0; b.x1 = b.y1 = -1;
            if (o.kind == K_CUTOUT) {        // ImageDraw.rectangle clips to the image
                b = bx[jj];
                if (b.x0 < 0) b.x0 = 0;
                if (b.y0 < 0) b.y0 = 0;
                if (b.x1 > P.W - 1) b.x1 = (int16_t)(P.W - 1);
                if (b.y1 > P.H - 1) b.y1 = (int16_t)(P.H - 1);
            }
            st.box[j] = b;
            st.suml[j] = 0ull;
        }
    }
    for (int i = threadIdx.x; i < 2 * 768; i += blockDim.x) (&st.hist[0][0])[i] = 0u;
}

__device__ __forceinline__ Ctx make_ctx(const AugParams& P, int idx, const ImgState& st) {
    Ctx c;
    c.raw = P.in + (size_t)idx * (size_t)P.H * (size_t)P.W * 3u;
    c.H = P.H; c.W = P.W;
    c.op[0] = st.op[0]; c.op[1] = st.op[1];
    c.box[0] = st.box[0]; c.box[1] = st.box[1];
    c.lut[0] = st.lut[0]; c.lut[1] = st.lut[1];
    return c;
}

// ---------------------------------------------------------------------------------------
// step 2: statistics of the image in front of slot L (0 or 1) over rows [y0, y1)
template <int L>
__device__ void accumulate_stats(const Ctx& c, int kind, int y0, int y1, uint32_t* hist, unsigned long long* suml) {
    const uint32_t n = (uint32_t)(y1 - y0) * (uint32_t)c.W;
    FastDiv dw; dw.init((uint32_t)c.W);
    if (kind == K_CONTRAST) {
        uint32_t local = 0;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t r = dw.div(i);
            local += luma_of(Level<L>::at(c, (int)(i - r * c.W), y0 + (int)r));
        }
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(suml, (unsigned long long)local);
    } else {
        const int lane = threadIdx.x & 31;
        for (uint32_t base = 0; base < n; base += blockDim.x) {      // warp-uniform trip count
            const uint32_t i = base + threadIdx.x;
            const bool valid = i < n;
            const uint32_t act = __ballot_sync(0xffffffffu, valid);
            if (!valid) continue;
            uint32_t r = dw.div(i);
            uint32_t p = Level<L>::at(c, (int)(i - r * c.W), y0 + (int)r);
            // warp-aggregated increments: lanes that hit the same bin elect one adder
            // (constant-colour regions would otherwise serialise 32-way on one bank)
            uint32_t b0 = p & 255u, b1 = 256u + ((p >> 8) & 255u), b2 = 512u + (p >> 16);
            uint32_t m0 = __match_any_sync(act, b0);
            if (lane == __ffs(m0) - 1) atomicAdd(&hist[b0], (uint32_t)__popc(m0));
            uint32_t m1 = __match_any_sync(act, b1);
            if (lane == __ffs(m1) - 1) atomicAdd(&hist[b1], (uint32_t)__popc(m1));
            uint32_t m2 = __match_any_sync(act, b2);
            if (lane == __ffs(m2) - 1) atomicAdd(&hist[b2], (uint32_t)__popc(m2));
        }
    }
}

// reduce slot j's partial statistics over the cluster and build its LUT
__device__ void build_slot_lut(const AugParams& P, ImgState& st, int j, cg::cluster_group& cluster) {
    const int kind = st.op[j].kind;
    const bool stats = kind_needs_hist(kind) || kind_needs_mean(kind);
    uint32_t mean = 0;
    if (stats) {
        if (P.bands > 1) cluster.sync(); else __syncthreads();     // partials complete everywhere
        if (kind_needs_hist(kind)) {
            for (int i = threadIdx.x; i < 768; i += blockDim.x) {
                uint32_t t = 0;
                for (int r = 0; r < P.bands; ++r) {
                    const uint32_t* rem = (P.bands > 1) ? cluster.map_shared_rank(&st.hist[j][0], r) : &st.hist[j][0];
                    t += rem[i];
                }
                st.tot[i] = t;
            }
        } else {
            unsigned long long t = 0;
            for (int r = 0; r < P.bands; ++r) {
                const unsigned long long* rem = (P.bands > 1) ? cluster.map_shared_rank(&st.suml[j], r) : &st.suml[j];
                t += *rem;
            }
            mean = contrast_mean(t, (uint32_t)P.H * (uint32_t)P.W);
        }
        __syncthreads();
    }
    if (kind_needs_hist(kind)) {
        const int t = threadIdx.x;
        if (t < 96) st.parts[t >> 5][t & 31] = hist_part(&st.tot[(t >> 5) * 256], t & 31);
        __syncthreads();
        if (t < 96)
            hist_lut_lane(kind, &st.tot[(t >> 5) * 256], st.parts[t >> 5], t & 31,
                          (uint32_t)P.H * (uint32_t)P.W, &st.lut[j][(t >> 5) * 256]);
    } else if (kind_uses_lut(kind)) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x)
            st.lut[j][i] = (uint8_t)lut_entry_static(st.op[j], (uint32_t)(i & 255), mean);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// step 3 helpers
__device__ __forceinline__ float normalise(const AugParams& P, int ch, uint32_t u) {
    return P.use_tab ? __ldg(P.norm_tab + ch * 256 + u) : fmaf((float)u, P.scale[ch], P.bias[ch]);
}

// four consecutive output pixels of row oy starting at ox0, as packed RGB; zmask bit k set
// when pixel k lies in the CutoutDefault zero box
__device__ __forceinline__ void quad_pixels(const AugParams& P, const Ctx& c, const Sample& s, bool fast,
                                            int ox0, int oy, uint32_t px[4], uint32_t& zmask) {
    zmask = 0;
    if (fast) {
        // pointwise-only program, 4-aligned source quad: three 32-bit loads of 12 contiguous bytes
        const int sx0 = (s.flip ? (P.out_w - 4 - ox0) : ox0) + s.crop_dx;
        const int ay = oy + s.crop_dy;
        uint32_t q[4] = {0u, 0u, 0u, 0u};
        const bool inside = (unsigned)sx0 < (unsigned)c.W && (unsigned)ay < (unsigned)c.H;
        if (inside) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(c.raw + ((size_t)ay * c.W + sx0) * 3u);
            uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            q[0] = w0 & 0xFFFFFFu;
            q[1] = (w0 >> 24) | ((w1 & 0xFFFFu) << 8);
            q[2] = (w1 >> 16) | ((w2 & 0xFFu) << 16);
            q[3] = w2 >> 8;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                q[k] = apply_pointwise(c, 1, apply_pointwise(c, 0, q[k], sx0 + k, ay), sx0 + k, ay);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            px[k] = s.flip ? q[3 - k] : q[k];
            const int ox = ox0 + k;
            if (P.use_zero_box && oy >= s.zero_box[0] && oy < s.zero_box[1] && ox >= s.zero_box[2] && ox < s.zero_box[3])
                zmask |= 1u << k;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ox = ox0 + k;
        px[k] = 0u;
        if (ox >= P.out_w) continue;
        int ax, ay; bool inside;
        if (!tail_source(s, P.use_zero_box != 0, P.out_w, c.H, c.W, ox, oy, ax, ay, inside)) { zmask |= 1u << k; continue; }
        if (inside) px[k] = Level<2>::at(c, ax, ay);
    }
}

template <int OUT> struct OutElem;
template <> struct OutElem<OUT_F16> { using T = __half; };
template <> struct OutElem<OUT_BF16> { using T = __nv_bfloat16; };
template <> struct OutElem<OUT_F32> { using T = float; };

template <int OUT>
__device__ __forceinline__ void store_plane4(void* out, size_t elem_off, const float v[4], bool vec, int nvalid) {
    using T = typename OutElem<OUT>::T;
    T* o = reinterpret_cast<T*>(out) + elem_off;
    if (vec) {
        if constexpr (OUT == OUT_F32) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else if constexpr (OUT == OUT_F16) {
            __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
            uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(o) = u;
        } else {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(o) = u;
        }
    } else {
        for (int k = 0; k < nvalid; ++k) {
            if constexpr (OUT == OUT_F32) o[k] = v[k];
            else if constexpr (OUT == OUT_F16) o[k] = __float2half_rn(v[k]);
            else o[k] = __float2bfloat16_rn(v[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------
template <int OUT, int NSRC>
__global__ void __launch_bounds__(kThreads) faa_augment_kernel(const __grid_constant__ AugParams P) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ ImgState st[NSRC];

    const int band = blockIdx.x;
    const int img = blockIdx.y;
    int src_idx[NSRC];
    src_idx[0] = P.first + img;
    if constexpr (NSRC == 2) src_idx[1] = P.partner[img];

    // ---- 1. decisions
#pragma unroll
    for (int s = 0; s < NSRC; ++s) load_program(P, src_idx[s], st[s]);
    __syncthreads();

    // ---- 2. statistics + LUTs, slot by slot (cluster-uniform control flow)
    const int y0 = (int)(((long long)band * P.H) / P.bands);
    const int y1 = (int)(((long long)(band + 1) * P.H) / P.bands);
    bool any_stats = false;
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kind = st[s].op[j].kind;
            if (kind_needs_hist(kind) || kind_needs_mean(kind)) {
                any_stats = true;
                Ctx c = make_ctx(P, src_idx[s], st[s]);
                if (j == 0) accumulate_stats<0>(c, kind, y0, y1, st[s].hist[0], &st[s].suml[0]);
                else        accumulate_stats<1>(c, kind, y0, y1, st[s].hist[1], &st[s].suml[1]);
            }
            if (kind_uses_lut(kind)) build_slot_lut(P, st[s], j, cluster);
        }
    }

    // ---- 3. output band
    const int oy0 = (int)(((long long)band * P.out_h) / P.bands);
    const int oy1 = (int)(((long long)(band + 1) * P.out_h) / P.bands);
    const uint32_t qpr = (uint32_t)(P.out_w + 3) >> 2;
    const uint32_t nq = (uint32_t)(oy1 - oy0) * qpr;
    FastDiv dq; dq.init(qpr);
    const bool vec = (P.out_w & 3) == 0;

    Ctx c0 = make_ctx(P, src_idx[0], st[0]);
    const Sample s0 = st[0].smp;
    const bool geom_ok = ((P.W & 3) == 0) && vec;
    const bool fast0 = geom_ok && kind_is_pointwise(c0.op[0].kind) && kind_is_pointwise(c0.op[1].kind) &&
                       ((s0.crop_dx & 3) == 0);
    Ctx c1; Sample s1; bool fast1 = false;
    if constexpr (NSRC == 2) {
        c1 = make_ctx(P, src_idx[1], st[1]);
        s1 = st[1].smp;
        fast1 = geom_ok && kind_is_pointwise(c1.op[0].kind) && kind_is_pointwise(c1.op[1].kind) && ((s1.crop_dx & 3) == 0);
    }

    for (uint32_t q = threadIdx.x; q < nq; q += blockDim.x) {
        const uint32_t r = dq.div(q);
        const int ox0 = (int)(q - r * qpr) * 4;
        const int oy = oy0 + (int)r;
        const int nvalid = min(4, P.out_w - ox0);
        uint32_t px[4], zmask;
        quad_pixels(P, c0, s0, fast0, ox0, oy, px, zmask);

        if constexpr (OUT == OUT_U8_HWC) {
            uint8_t* o = reinterpret_cast<uint8_t*>(P.out) + (((size_t)img * P.out_h + oy) * P.out_w + ox0) * 3u;
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((zmask >> k) & 1u) px[k] = 0u;
            if (vec) {
                uint32_t* w = reinterpret_cast<uint32_t*>(o);
                w[0] = px[0] | (px[1] << 24);
                w[1] = (px[1] >> 8) | (px[2] << 16);
                w[2] = (px[2] >> 16) | (px[3] << 8);
            } else {
                for (int k = 0; k < nvalid; ++k) {
                    o[3 * k] = (uint8_t)px[k]; o[3 * k + 1] = (uint8_t)(px[k] >> 8); o[3 * k + 2] = (uint8_t)(px[k] >> 16);
                }
            }
        } else {
            uint32_t px1[4], zmask1 = 0;
            if constexpr (NSRC == 2) quad_pixels(P, c1, s1, fast1, ox0, oy, px1, zmask1);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a = ((zmask >> k) & 1u) ? 0.0f : normalise(P, ch, (px[k] >> (8 * ch)) & 255u);
                    if constexpr (NSRC == 2) {
                        float b = ((zmask1 >> k) & 1u) ? 0.0f : normalise(P, ch, (px1[k] >> (8 * ch)) & 255u);
                        a = f_add(f_mul(a, P.lam), f_mul(b, P.one_minus_lam));       // aug_mixup.py:21
                    }
                    v[k] = a;
                }
                const size_t off = (((size_t)img * 3 + ch) * P.out_h + oy) * (size_t)P.out_w + ox0;
                store_plane4<OUT>(P.out, off, v, vec, nvalid);
            }
        }
    }

    // a CTA must not exit while cluster peers may still read its partial statistics
    if (any_stats && P.bands > 1) cluster.sync();
}

// ---------------------------------------------------------------------------------------
__global__ void faa_philox_kernel(const __grid_constant__ PhiloxParams P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    Sample s; Box bx[8];
    philox_sample(P.rng, P.rng.first_index + (uint64_t)i, P.ops, P.probs, P.n_sub, P.n_op, P.H, P.W,
                  P.out_h, P.out_w, s, bx);
    P.samples[i] = s;
    for (int j = 0; j < P.n_op; ++j) P.boxes[(size_t)i * P.n_op + j] = bx[j];
}

// out[i] = data[i]*lam + data[perm[i]]*(1-lam), fp32 math (aug_mixup.py:13-23)
template <typename T>
__global__ void faa_mixup_kernel(const T* __restrict__ data, T* __restrict__ out, const int64_t* __restrict__ perm,
                                 int64_t n_per, float lam, float oml) {
    const int b = blockIdx.y;
    const T* a = data + (size_t)b * n_per;
    const T* p = data + (size_t)perm[b] * n_per;
    T* o = out + (size_t)b * n_per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += (int64_t)gridDim.x * blockDim.x) {
        float x = (float)a[i], y = (float)p[i];
        o[i] = (T)f_add(f_mul(x, lam), f_mul(y, oml));
    }
}

// ---------------------------------------------------------------------------------------
int pick_bands(int H, int W, int out_h, int out_w) {
    // aim for >= ~1024 output quads per CTA; cluster size must be a power of two <= 8
    long long quads = (long long)out_h * ((out_w + 3) / 4);
    int b = 1;
    while (b < 8 && quads / (b * 2) >= 1024 && b * 2 <= H && b * 2 <= out_h) b *= 2;
    return b;
}

template <int OUT, int NSRC>
static cudaError_t launch_one(const AugParams& p, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.bands, (unsigned)p.B, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.bands;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, faa_augment_kernel<OUT, NSRC>, p);
}

cudaError_t launch_augment(const AugParams& p, int out_type, cudaStream_t stream) {
    if (p.B <= 0) return cudaSuccess;
    const bool mix = p.partner != nullptr;
    switch (out_type) {
    case OUT_F16:  return mix ? launch_one<OUT_F16, 2>(p, stream)  : launch_one<OUT_F16, 1>(p, stream);
    case OUT_BF16: return mix ? launch_one<OUT_BF16, 2>(p, stream) : launch_one<OUT_BF16, 1>(p, stream);
    case OUT_F32:  return mix ? launch_one<OUT_F32, 2>(p, stream)  : launch_one<OUT_F32, 1>(p, stream);
    case OUT_U8_HWC: return launch_one<OUT_U8_HWC, 1>(p, stream);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_philox(const PhiloxParams& p, cudaStream_t stream) {
    if (p.B <= 0) return cudaSuccess;
    faa_philox_kernel<<<(p.B + 127) / 128, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_mixup(const void* data, void* out, const int64_t* perm, int batch, int64_t n_per_sample,
                         int dtype, float lam, float oml, cudaStream_t stream) {
    if (batch <= 0 || n_per_sample <= 0) return cudaSuccess;
    unsigned gx = (unsigned)((n_per_sample + 255) / 256);
    if (gx > 64) gx = 64;
    dim3 grid(gx, (unsigned)batch, 1);
    switch (dtype) {
    case OUT_F16:  faa_mixup_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)data, (__half*)out, perm, n_per_sample, lam, oml); break;
    case OUT_BF16: faa_mixup_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)data, (__nv_bfloat16*)out, perm, n_per_sample, lam, oml); break;
    case OUT_F32:  faa_mixup_kernel<float><<<grid, 256, 0, stream>>>((const float*)data, (float*)out, perm, n_per_sample, lam, oml); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace faa
