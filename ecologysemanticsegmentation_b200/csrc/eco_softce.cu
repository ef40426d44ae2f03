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
        partials[(int64_t)blockIdx.x * 2 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp < 2) {
        double v = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(partials + (int64_t)i * 2 + warp);
        v = warp_sum(v);
        if (lane == 0) sums_out[warp] = v;
    }
    if (threadIdx.x == 0) *counter = 0;
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
    ECO_SCE(softce_grad_kernel, a->dtype, b->dtype, <<<grid, kSceThreads, 0, st>>>(g, upstream));
    return check_cuda(cudaGetLastError(), "softce_grad_kernel launch");
}
