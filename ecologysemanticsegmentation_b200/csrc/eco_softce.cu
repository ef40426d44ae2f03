// Soft-label cross entropy over the channel dim (ess/loss_functions.py:29-30):
//   F.cross_entropy(pred, gt) + bw * F.cross_entropy(1 - pred, 1 - gt), float targets of pred's shape.
// On the 1-channel slices every losses_fn feeds it this is identically 0 (log_softmax over one channel);
// this kernel serves the stand-alone multi-channel call.  One thread per pixel, coalesced along H*W,
// online log-sum-exp over the C channel planes.
#include "eco_common.cuh"

namespace eco {

struct SceArgs {
    const void* a;
    const void* b;
    int64_t a_sn, a_sc, b_sn, b_sc;
    int32_t N, C;
    int64_t HW;
    int32_t need_bg;
};

template <typename T>
__device__ __forceinline__ float ld1(const void* base, int64_t idx) {
    return Vec4<T>::load1(reinterpret_cast<const T*>(base) + idx);
}

// sum_c a_c * log_softmax(b)_c for one pixel; `flip` evaluates it on (1-a, 1-b)
template <typename TA, typename TB>
__device__ __forceinline__ float pixel_soft_ce(const SceArgs& p, int64_t ao, int64_t bo, bool flip, float* lse_out,
                                               float* asum_out) {
    float m = -INFINITY, s = 0.f, ab = 0.f, as = 0.f;
    for (int c = 0; c < p.C; ++c) {
        float a = ld1<TA>(p.a, ao + c * p.a_sc), b = ld1<TB>(p.b, bo + c * p.b_sc);
        if (flip) { a = 1.0f - a; b = 1.0f - b; }
        const float mn = fmaxf(m, b);
        s = s * expf(m - mn) + expf(b - mn);
        m = mn;
        ab = fmaf(a, b, ab);
        as += a;
    }
    const float lse = m + logf(s);
    if (lse_out) *lse_out = lse;
    if (asum_out) *asum_out = as;
    return ab - lse * as;
}

constexpr int kSceThreads = 256;
constexpr int kSceMaxCtas = 148 * 8;

// CTA reduction of the two sums, CTA partial -> global, the last CTA adds the partials in a fixed order and re-arms the counter
__device__ __forceinline__ void sce_finish(double acc0, double acc1, unsigned int* __restrict__ counter,
                                           double* __restrict__ partials, double* __restrict__ sums_out, int cta, int n_ctas) {
    __shared__ double sm[kSceThreads / 32][2];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    if (lane == 0) { sm[warp][0] = acc0; sm[warp][1] = acc1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int w = 0; w < kSceThreads / 32; ++w) v += sm[w][threadIdx.x];
        partials[(int64_t)cta * 2 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == (unsigned int)n_ctas - 1u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp < 2) {
        double v = 0.0;
        for (int i = lane; i < n_ctas; i += 32) v += __ldcg(partials + (int64_t)i * 2 + warp);
        v = warp_sum(v);
        if (lane == 0) sums_out[warp] = v;
    }
    if (threadIdx.x == 0) *counter = 0;
}

// ---------------------------------------------------------------------------------------------
// The usual shapes (2..4 channels, 16-byte aligned planes): four consecutive pixels per thread, every plane read with ONE
// 128-bit load per thread and kept in registers -- the gradient kernel reads its input once instead of twice --, MUFU
// exp2 / log2 (1e-7 relative, the tolerance is 1e-5), softmax from the exponentials of the log-sum-exp instead of a second
// exp per channel, no 64-bit division per pixel (grid.y walks the images).
// ---------------------------------------------------------------------------------------------
constexpr float kSceLog2e = 1.4426950408889634f, kSceLn2 = 0.6931471805599453f;

// one pixel: e[c] = exp(b_c - max), returns the log-sum-exp; FLIP evaluates on (1 - a, 1 - b)
template <int C, bool FLIP>
__device__ __forceinline__ float sce_pixel(const float (&a)[C], const float (&b)[C], float (&e)[C], float& s, float& ab, float& as) {
    float bb[C], aa[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { bb[c] = FLIP ? 1.0f - b[c] : b[c]; aa[c] = FLIP ? 1.0f - a[c] : a[c]; }
    float m = bb[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, bb[c]);
    s = 0.f; ab = 0.f; as = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        e[c] = ex2_approx((bb[c] - m) * kSceLog2e);
        s += e[c];
        ab = fmaf(aa[c], bb[c], ab);
        as += aa[c];
    }
    return m + lg2_approx(s) * kSceLn2;
}

template <typename TA, typename TB, int C>
__global__ void __launch_bounds__(kSceThreads)
softce_stats_vec_kernel(SceArgs p, unsigned int* __restrict__ counter, double* __restrict__ partials,
                        double* __restrict__ sums_out) {
    const int64_t units = p.HW / 4;
    double acc0 = 0.0, acc1 = 0.0;
    for (int n = blockIdx.y; n < p.N; n += gridDim.y) {
        const TA* ap = reinterpret_cast<const TA*>(p.a) + (int64_t)n * p.a_sn;
        const TB* bp = reinterpret_cast<const TB*>(p.b) + (int64_t)n * p.b_sn;
        for (int64_t u = (int64_t)blockIdx.x * kSceThreads + threadIdx.x; u < units; u += (int64_t)gridDim.x * kSceThreads) {
            float a[C][4], b[C][4];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                Vec4<TA>::load(ap + c * p.a_sc + 4 * u, a[c]);
                Vec4<TB>::load(bp + c * p.b_sc + 4 * u, b[c]);
            }
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float av[C], bv[C], e[C], s, ab, as;
#pragma unroll
                for (int c = 0; c < C; ++c) { av[c] = a[c][v]; bv[c] = b[c][v]; }
                const float lse = sce_pixel<C, false>(av, bv, e, s, ab, as);
                t0 += ab - lse * as;
                if (p.need_bg) {
                    const float lse_f = sce_pixel<C, true>(av, bv, e, s, ab, as);
                    t1 += ab - lse_f * as;
                }
            }
            acc0 += (double)t0;
            acc1 += (double)t1;
        }
    }
    sce_finish(acc0, acc1, counter, partials, sums_out, (int)(blockIdx.y * gridDim.x + blockIdx.x), (int)(gridDim.x * gridDim.y));
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(kSceThreads)
softce_stats_kernel(SceArgs p, unsigned int* __restrict__ counter, double* __restrict__ partials,
                    double* __restrict__ sums_out) {
    const int64_t total = (int64_t)p.N * p.HW;
    double acc0 = 0.0, acc1 = 0.0;
    for (int64_t q = (int64_t)blockIdx.x * kSceThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kSceThreads) {
        const int64_t n = q / p.HW, e = q - n * p.HW;
        const int64_t ao = n * p.a_sn + e, bo = n * p.b_sn + e;
        acc0 += (double)pixel_soft_ce<TA, TB>(p, ao, bo, false, nullptr, nullptr);
        if (p.need_bg) acc1 += (double)pixel_soft_ce<TA, TB>(p, ao, bo, true, nullptr, nullptr);
    }
    sce_finish(acc0, acc1, counter, partials, sums_out, (int)blockIdx.x, (int)gridDim.x);
}

struct SceGradArgs {
    SceArgs p;
    void* ga;
    void* gb;
    int64_t ga_sn, ga_sc, gb_sn, gb_sc;
    float bw;
    float inv_npix;
};

template <typename TA, typename TB>
__global__ void __launch_bounds__(kSceThreads)
softce_grad_kernel(SceGradArgs g, const float* __restrict__ upstream) {
    const SceArgs& p = g.p;
    const int64_t total = (int64_t)p.N * p.HW;
    const float w = upstream[0] * g.inv_npix;
    for (int64_t q = (int64_t)blockIdx.x * kSceThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kSceThreads) {
        const int64_t n = q / p.HW, e = q - n * p.HW;
        const int64_t ao = n * p.a_sn + e, bo = n * p.b_sn + e;
        float lse, as, lse_f = 0.f, as_f = 0.f;
        pixel_soft_ce<TA, TB>(p, ao, bo, false, &lse, &as);
        if (p.need_bg) pixel_soft_ce<TA, TB>(p, ao, bo, true, &lse_f, &as_f);
        for (int c = 0; c < p.C; ++c) {
            const float a = ld1<TA>(p.a, ao + c * p.a_sc), b = ld1<TB>(p.b, bo + c * p.b_sc);
            const float ls = b - lse;
            float da = -ls, db = expf(ls) * as - a;
            if (p.need_bg) {
                const float lsf = (1.0f - b) - lse_f;
                da += g.bw * lsf;
                db -= g.bw * (expf(lsf) * as_f - (1.0f - a));
            }
            if (g.ga) Vec4<TA>::store1(reinterpret_cast<TA*>(g.ga) + n * g.ga_sn + c * g.ga_sc + e, w * da);
            if (g.gb) Vec4<TB>::store1(reinterpret_cast<TB*>(g.gb) + n * g.gb_sn + c * g.gb_sc + e, w * db);
        }
    }
}

template <typename TA, typename TB, int C>
__global__ void __launch_bounds__(kSceThreads)
softce_grad_vec_kernel(SceGradArgs g, const float* __restrict__ upstream) {
    const SceArgs& p = g.p;
    const int64_t units = p.HW / 4;
    const float w = upstream[0] * g.inv_npix;
    for (int n = blockIdx.y; n < p.N; n += gridDim.y) {
        const TA* ap = reinterpret_cast<const TA*>(p.a) + (int64_t)n * p.a_sn;
        const TB* bp = reinterpret_cast<const TB*>(p.b) + (int64_t)n * p.b_sn;
        for (int64_t u = (int64_t)blockIdx.x * kSceThreads + threadIdx.x; u < units; u += (int64_t)gridDim.x * kSceThreads) {
            float a[C][4], b[C][4], da[C][4], db[C][4];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                Vec4<TA>::load(ap + c * p.a_sc + 4 * u, a[c]);
                Vec4<TB>::load(bp + c * p.b_sc + 4 * u, b[c]);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float av[C], bv[C], e[C], s, ab, as;
#pragma unroll
                for (int c = 0; c < C; ++c) { av[c] = a[c][v]; bv[c] = b[c][v]; }
                const float lse = sce_pixel<C, false>(av, bv, e, s, ab, as);
                const float r = as * rcp_approx(s);                  // softmax_c * sum a = e_c * r
#pragma unroll
                for (int c = 0; c < C; ++c) { da[c][v] = lse - bv[c]; db[c][v] = fmaf(e[c], r, -av[c]); }
                if (p.need_bg) {
                    const float lse_f = sce_pixel<C, true>(av, bv, e, s, ab, as);
                    const float rf = as * rcp_approx(s);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        da[c][v] = fmaf(g.bw, (1.0f - bv[c]) - lse_f, da[c][v]);
                        db[c][v] -= g.bw * fmaf(e[c], rf, -(1.0f - av[c]));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int v = 0; v < 4; ++v) { da[c][v] *= w; db[c][v] *= w; }
                if (g.ga) Vec4<TA>::store(reinterpret_cast<TA*>(g.ga) + (int64_t)n * g.ga_sn + c * g.ga_sc + 4 * u, da[c]);
                if (g.gb) Vec4<TB>::store(reinterpret_cast<TB*>(g.gb) + (int64_t)n * g.gb_sn + c * g.gb_sc + 4 * u, db[c]);
            }
        }
    }
}

// planes that the 128-bit kernels can walk: base pointer, image and channel strides and H*W all multiples of four elements
// (16 bytes for float32, 8 for bf16)
static bool sce_vec_ok(const void* ptr, int64_t sn, int64_t sc, int dtype, int64_t HW) {
    const int64_t esz = dtype == ECO_BF16 ? 2 : 4;
    return reinterpret_cast<uintptr_t>(ptr) % (4 * esz) == 0 && sn % 4 == 0 && sc % 4 == 0 && HW % 4 == 0;
}
// grid of the 128-bit kernels: x over the 4-pixel units of an image, y over the images, at most kSceMaxCtas CTAs
static dim3 sce_vec_grid(int device, int32_t N, int64_t HW) {
    const int sms = sm_count_cached(device);
    const int64_t cap = (int64_t)(sms > 0 ? sms : 148) * 8 < kSceMaxCtas ? (int64_t)(sms > 0 ? sms : 148) * 8 : kSceMaxCtas;
    int64_t gx = (HW / 4 + kSceThreads - 1) / kSceThreads;
    if (gx < 1) gx = 1;
    if (gx > cap) gx = cap;
    int64_t gy = cap / gx;
    if (gy < 1) gy = 1;
    if (gy > N) gy = N;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

#define ECO_SCE_VEC(KERNEL, adt, bdt, CC, ...)                                                      \
    do {                                                                                             \
        if (adt == ECO_F32 && bdt == ECO_F32) KERNEL<float, float, CC> __VA_ARGS__;                  \
        else if (adt == ECO_BF16 && bdt == ECO_F32) KERNEL<__nv_bfloat16, float, CC> __VA_ARGS__;    \
        else if (adt == ECO_F32 && bdt == ECO_BF16) KERNEL<float, __nv_bfloat16, CC> __VA_ARGS__;    \
        else KERNEL<__nv_bfloat16, __nv_bfloat16, CC> __VA_ARGS__;                                   \
    } while (0)

static int sce_fill(SceArgs& p, const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, int need_bg) {
    if (N <= 0 || C <= 0 || HW <= 0) { set_error("empty input"); return -2; }
    if (!a || !b || !a->ptr || !b->ptr) { set_error("null input view"); return -1; }
    p.a = a->ptr; p.b = b->ptr; p.a_sn = a->sn; p.a_sc = a->sc; p.b_sn = b->sn; p.b_sc = b->sc;
    p.N = N; p.C = C; p.HW = HW; p.need_bg = need_bg;
    return 0;
}
static int sce_grid(int device, int64_t total) {
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -1;
    int64_t g = (total + kSceThreads - 1) / kSceThreads;
    const int64_t cap = (int64_t)sms * 8 < kSceMaxCtas ? (int64_t)sms * 8 : kSceMaxCtas;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

#define ECO_SCE(KERNEL, adt, bdt, ...)                                                          \
    do {                                                                                         \
        if (adt == ECO_F32 && bdt == ECO_F32) KERNEL<float, float> __VA_ARGS__;                  \
        else if (adt == ECO_BF16 && bdt == ECO_F32) KERNEL<__nv_bfloat16, float> __VA_ARGS__;    \
        else if (adt == ECO_F32 && bdt == ECO_BF16) KERNEL<float, __nv_bfloat16> __VA_ARGS__;    \
        else KERNEL<__nv_bfloat16, __nv_bfloat16> __VA_ARGS__;                                   \
    } while (0)

}  // namespace eco

using namespace eco;

extern "C" int64_t eco_softce_ws_bytes(void) { return 256 + (int64_t)kSceMaxCtas * 2 * (int64_t)sizeof(double); }

extern "C" int eco_softce_stats(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, int32_t need_bg,
                                void* ws, int64_t ws_bytes, double* sums_out, int device, void* stream) {
    SceArgs p{};
    int rc = sce_fill(p, a, b, N, C, HW, need_bg);
    if (rc) return rc;
    if (!ws || ws_bytes < eco_softce_ws_bytes() || !sums_out) { set_error("workspace too small or null output"); return -5; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const int grid = sce_grid(device, (int64_t)N * HW);
    if (grid < 0) return -10;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C >= 2 && C <= 4 && sce_vec_ok(a->ptr, a->sn, a->sc, a->dtype, HW) && sce_vec_ok(b->ptr, b->sn, b->sc, b->dtype, HW)) {
        const dim3 vg = sce_vec_grid(device, N, HW);
        if (C == 2) ECO_SCE_VEC(softce_stats_vec_kernel, a->dtype, b->dtype, 2, <<<vg, kSceThreads, 0, st>>>(p, counter, partials, sums_out));
        else if (C == 3) ECO_SCE_VEC(softce_stats_vec_kernel, a->dtype, b->dtype, 3, <<<vg, kSceThreads, 0, st>>>(p, counter, partials, sums_out));
        else ECO_SCE_VEC(softce_stats_vec_kernel, a->dtype, b->dtype, 4, <<<vg, kSceThreads, 0, st>>>(p, counter, partials, sums_out));
        return check_cuda(cudaGetLastError(), "softce_stats_vec_kernel launch");
    }
    ECO_SCE(softce_stats_kernel, a->dtype, b->dtype, <<<grid, kSceThreads, 0, st>>>(p, counter, partials, sums_out));
    return check_cuda(cudaGetLastError(), "softce_stats_kernel launch");
}

extern "C" int eco_softce_grad(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW,
                               double background_weight, double n_pix_total, const float* upstream, const EcoOut* ga,
                               const EcoOut* gb, int device, void* stream) {
    SceGradArgs g{};
    int rc = sce_fill(g.p, a, b, N, C, HW, background_weight != 0.0);
    if (rc) return rc;
    if (!upstream) { set_error("null upstream"); return -5; }
    const bool want_a = ga && ga->ptr, want_b = gb && gb->ptr;
    if (!want_a && !want_b) return 0;
    if ((want_a && ga->dtype != a->dtype) || (want_b && gb->dtype != b->dtype)) { set_error("gradient dtype must match its slot"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    if (want_a) { g.ga = ga->ptr; g.ga_sn = ga->sn; g.ga_sc = ga->sc; }
    if (want_b) { g.gb = gb->ptr; g.gb_sn = gb->sn; g.gb_sc = gb->sc; }
    g.bw = (float)background_weight;
    g.inv_npix = (float)(1.0 / n_pix_total);
    const int grid = sce_grid(device, (int64_t)N * HW);
    if (grid < 0) return -10;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C >= 2 && C <= 4 && sce_vec_ok(a->ptr, a->sn, a->sc, a->dtype, HW) && sce_vec_ok(b->ptr, b->sn, b->sc, b->dtype, HW) &&
        (!want_a || sce_vec_ok(ga->ptr, ga->sn, ga->sc, ga->dtype, HW)) && (!want_b || sce_vec_ok(gb->ptr, gb->sn, gb->sc, gb->dtype, HW))) {
        const dim3 vg = sce_vec_grid(device, N, HW);
        if (C == 2) ECO_SCE_VEC(softce_grad_vec_kernel, a->dtype, b->dtype, 2, <<<vg, kSceThreads, 0, st>>>(g, upstream));
        else if (C == 3) ECO_SCE_VEC(softce_grad_vec_kernel, a->dtype, b->dtype, 3, <<<vg, kSceThreads, 0, st>>>(g, upstream));
        else ECO_SCE_VEC(softce_grad_vec_kernel, a->dtype, b->dtype, 4, <<<vg, kSceThreads, 0, st>>>(g, upstream));
        return check_cuda(cudaGetLastError(), "softce_grad_vec_kernel launch");
    }
    ECO_SCE(softce_grad_kernel, a->dtype, b->dtype, <<<grid, kSceThreads, 0, st>>>(g, upstream));
    return check_cuda(cudaGetLastError(), "softce_grad_kernel launch");
}
