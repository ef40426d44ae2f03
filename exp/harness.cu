// Experiment harness (not part of the library): launches the second-generation composite kernels directly, prints the
// in-kernel %globaltimer timeline of the fused step (ECO_V2_TIMELINE), and times them next to the C ABI entry points
// (which, for fp32 logits, route to the same kernels -- the "shipped" columns are a launch-path check, not a baseline).
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a exp/harness.cu \
//        ecologysemanticsegmentation_b200/csrc/build/eco_api.o -o exp/harness
#define ECO_V2_TIMELINE 1
#include "../ecologysemanticsegmentation_b200/csrc/eco_composite.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define RC(x) do { int r_ = (x); if (r_) { printf("rc=%d (%s) at %s:%d\n", r_, eco_last_error(), __FILE__, __LINE__); exit(1); } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline double urand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (rng_state >> 11) * (1.0 / 9007199254740992.0); }
static inline float nrand() { double u = urand() + 1e-12, v = urand(); return (float)(sqrt(-2 * log(u)) * cos(6.283185307179586 * v)); }

template <typename F>
static float time_us(F f, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) f(i);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f(i);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / iters;
}

int main(int argc, char** argv) {
    const int N = 54, C = 3, S = argc > 1 ? atoi(argv[1]) : 256;
    const int64_t HW = (int64_t)S * S, E = (int64_t)N * C * HW;
    const int NSETS = 4;
    std::vector<float> hz(E), hg(E);
    for (int64_t n = 0; n < N; ++n)
        for (int64_t i = 0; i < HW; ++i) {
            const double u = urand();
            for (int c = 0; c < 3; ++c) hz[(n * C + c) * HW + i] = nrand();
            hg[(n * C + 0) * HW + i] = u < 0.5;
            hg[(n * C + 1) * HW + i] = u < 0.5 * 0.43197708;
            hg[(n * C + 2) * HW + i] = u < 0.5 * 0.22319692;
        }
    // a few exact ties and saturated logits
    for (int k = 0; k < 64; ++k) { int64_t i = (int64_t)(urand() * HW); hz[(0 * C + 1) * HW + i] = hz[(0 * C + 0) * HW + i]; }
    for (int k = 0; k < 64; ++k) { int64_t i = (int64_t)(urand() * HW); hz[(1 * C + 2) * HW + i] = 9.f; hz[(1 * C + 0) * HW + i] = -9.f; }
    float *z[NSETS], *g[NSETS], *o_old, *o_new;
    for (int k = 0; k < NSETS; ++k) {
        CK(cudaMalloc(&z[k], E * 4)); CK(cudaMalloc(&g[k], E * 4));
        CK(cudaMemcpy(z[k], hz.data(), E * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(g[k], hg.data(), E * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&o_old, E * 4)); CK(cudaMalloc(&o_new, E * 4));
    void* ws; const int64_t wsb = eco_composite3_ws_bytes(); CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0, wsb));
    double *acc, *jac, *scale_dev; float *losses, *up;
    CK(cudaMalloc(&acc, 128 * 8)); CK(cudaMalloc(&jac, 21 * 7 * 7 * 8)); CK(cudaMalloc(&scale_dev, 21 * 8));
    CK(cudaMalloc(&losses, 7 * 4)); CK(cudaMalloc(&up, 7 * 4));
    const double r[3] = {1., 0.43197708, 0.22319692};
    double sc[21] = {2, 2, 2};
    { int t = 3; for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) { const double wi = 1 / r[i], wj = 1 / r[j], wd = 1 / (r[i] - r[j]);
        sc[t++] = 2 * wj; sc[t++] = 2 * wi; sc[t++] = 2 * wd; sc[t++] = 2 * wi; sc[t++] = 2 * wd; sc[t++] = 2 * wi * wi * wj; } }
    CK(cudaMemcpy(scale_dev, sc, sizeof(sc), cudaMemcpyHostToDevice));
    const float hup[7] = {0, 1, 0, 0, 1, 1, 1};
    CK(cudaMemcpy(up, hup, sizeof(hup), cudaMemcpyHostToDevice));

    auto view = [&](const float* p) { EcoView v{}; v.ptr = p; v.sn = C * HW; v.sc = HW; v.dtype = ECO_F32; return v; };
    auto outv = [&](float* p) { EcoOut v{}; v.ptr = p; v.sn = C * HW; v.sc = HW; v.dtype = ECO_F32; return v; };
    EcoView vz[NSETS], vg[NSETS];
    for (int k = 0; k < NSETS; ++k) { vz[k] = view(z[k]); vg[k] = view(g[k]); }
    EcoOut oo = outv(o_old), on = outv(o_new);

    RC(eco_composite3_stats(&vz[0], &vg[0], N, HW, 1, ws, wsb, acc, 0, nullptr));
    RC(eco_composite3_finalize(acc, nullptr, scale_dev, losses, jac, nullptr, 0, nullptr));
    RC(eco_composite3_grad(&vz[0], &vg[0], N, HW, 1, jac, up, &oo, 0, nullptr));
    CK(cudaDeviceSynchronize());
    float hl[7]; CK(cudaMemcpy(hl, losses, 28, cudaMemcpyDeviceToHost));
    printf("losses:"); for (int k = 0; k < 7; ++k) printf(" %.6f", hl[k]); printf("\n");

    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    using namespace eco::v2;
    CK(cudaFuncSetAttribute(composite3_grad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    CK(cudaFuncSetAttribute(composite3_fused_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    double* acc_glob = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    double* partials = acc_glob + 128;
    float *o_f, *losses2; CK(cudaMalloc(&o_f, E * 4)); CK(cudaMalloc(&losses2, 28));
    auto cmp = [&](const char* what, const float* ref_d, const float* new_d) {
        std::vector<float> ho(E), hn(E);
        CK(cudaMemcpy(ho.data(), ref_d, E * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hn.data(), new_d, E * 4, cudaMemcpyDeviceToHost));
        double mx = 0, md = 0, l2 = 0, l2d = 0; int64_t worst = 0;
        for (int64_t i = 0; i < E; ++i) {
            const double d = fabs((double)ho[i] - hn[i]);
            if (d > md || d != d) { md = d; worst = i; }
            mx = fmax(mx, fabs((double)ho[i])); l2 += (double)ho[i] * ho[i]; l2d += d * d;
        }
        printf("%s: max|d|/max|ref| = %.3e   relL2 = %.3e   (worst idx %lld: %.9g vs %.9g)\n", what, md / mx, sqrt(l2d / l2), (long long)worst, ho[worst], hn[worst]);
    };
    auto launch_grad = [&](int k, float* out) {
        CompGradArgs ga{}; fill_comp(ga.a, &vz[k], &vg[k], N, HW, 4);
        ga.gx = out; ga.gx_sn = C * HW; ga.gx_sc = HW;
        composite3_grad_v2_kernel<<<sms, kThreads, kSmemBytes>>>(ga, jac, up);
    };
    eco::v2::V2Ws* ws2; CK(cudaMalloc(&ws2, sizeof(eco::v2::V2Ws))); CK(cudaMemset(ws2, 0, sizeof(eco::v2::V2Ws)));
    auto launch_fused = [&](int k, float* out, float* lo) {
        CompGradArgs ga{}; fill_comp(ga.a, &vz[k], &vg[k], N, HW, 4);
        ga.gx = out; ga.gx_sn = C * HW; ga.gx_sc = HW;
        XchArgs xch{}; xch.world = 1; xch.status = counter + 32;
        const double* sd = scale_dev; const float* u = up;
        void* args[] = {&ga, (void*)&sd, (void*)&u, &ws2, &acc_glob, &lo, &xch};
        CK(cudaLaunchCooperativeKernel((const void*)composite3_fused_v2_kernel, dim3(sms), dim3(kThreads), args, kSmemBytes, nullptr));
    };
    launch_grad(0, o_new); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    cmp("grad v2 vs shipped grad", o_old, o_new);
    // shipped fused as the reference for losses + gradient
    RC(eco_composite3_fused(&vz[0], &vg[0], N, HW, 1, scale_dev, up, ws, wsb, losses, &oo, 0, nullptr));
    CK(cudaDeviceSynchronize());
    launch_fused(0, o_f, losses2); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    cmp("fused v2 vs shipped fused", o_old, o_f);
    float h2[7]; CK(cudaMemcpy(hl, losses, 28, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, losses2, 28, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 7; ++k) printf("  loss[%d] shipped %.7g  v2 %.7g  rel %.2e\n", k, hl[k], h2[k], fabs(hl[k] - h2[k]) / fmax(1e-30, fabs(hl[k])));
    // all four gradient variants against the shipped two-kernel path
    for (int var = 0; var < 4; ++var) {
        const float hv[7] = {0, (var & 1) ? 1.f : 0.f, (var & 2) ? 1.f : 0.f, 0.5f, 1, 1, 1};
        CK(cudaMemcpy(up, hv, sizeof(hv), cudaMemcpyHostToDevice));
        RC(eco_composite3_grad(&vz[0], &vg[0], N, HW, 1, jac, up, &oo, 0, nullptr));
        launch_grad(0, o_new); CK(cudaDeviceSynchronize());
        char nm[64]; snprintf(nm, 64, "grad v2 variant sig=%d fl=%d", var & 1, (var >> 1) & 1);
        cmp(nm, o_old, o_new);
        RC(eco_composite3_fused(&vz[0], &vg[0], N, HW, 1, scale_dev, up, ws, wsb, losses, &oo, 0, nullptr));
        launch_fused(0, o_f, losses2); CK(cudaDeviceSynchronize());
        snprintf(nm, 64, "fused v2 variant sig=%d fl=%d", var & 1, (var >> 1) & 1);
        cmp(nm, o_old, o_f);
    }
    CK(cudaMemcpy(up, hup, sizeof(hup), cudaMemcpyHostToDevice));
    const int iters = 100;
    float t_old = time_us([&](int i) { eco_composite3_grad(&vz[i % NSETS], &vg[i % NSETS], N, HW, 1, jac, up, &oo, 0, nullptr); }, iters);
    float t_st = time_us([&](int i) { eco_composite3_stats(&vz[i % NSETS], &vg[i % NSETS], N, HW, 1, ws, wsb, acc, 0, nullptr); }, iters);
    float t_fu = time_us([&](int i) { eco_composite3_fused(&vz[i % NSETS], &vg[i % NSETS], N, HW, 1, scale_dev, up, ws, wsb, losses, &oo, 0, nullptr); }, iters);
    float t_g2 = time_us([&](int i) { launch_grad(i % NSETS, o_new); }, iters);
    float t_f2 = time_us([&](int i) { launch_fused(i % NSETS, o_f, losses2); }, iters);
    {
        { static unsigned long long z16[1024*16]; CK(cudaMemcpyToSymbol(eco::v2::g_timeline, z16, sizeof(z16))); }
        launch_fused(0, o_f, losses2); CK(cudaDeviceSynchronize());
        static unsigned long long tl[1024 * 16];
        CK(cudaMemcpyFromSymbol(tl, eco::v2::g_timeline, sizeof(tl)));
        unsigned long long t0 = ~0ull; for (int b = 0; b < sms; ++b) t0 = tl[b * 16] < t0 ? tl[b * 16] : t0;
        const int order[14] = {0, 1, 7, 8, 9, 10, 2, 3, 11, 12, 13, 4, 5, 6};
        const char* nm[14] = {"start", "pass1 loop end", "end", "closed forms end", "pass2 loop end", "end", "end", "sums->layout", "partial stored+fence", "arrival atomic", "last: loads done", "cf: sums loaded", "cf: rows done", "cf: make_coef done"};
        const char* nm2[14] = {"start", "pass1 loop end", "stats_finish end", "flag seen", "coef ready", "pass2 loop end", "end", "sums->layout", "partial stored+fence", "arrival atomic", "last: loads done", "cf: sums loaded", "cf: rows done", "cf: make_coef done"};
        (void)nm;
        for (int b = 0; b < 2; ++b) { printf("  cta %d rows (start,end):", b); for (int k = 0; k < 7; ++k) printf(" k%d %.2f-%.2f", k, (double)(tl[1024*16-64+b*16+k]-t0)*1e-3, (double)(tl[1024*16-64+b*16+8+k]-t0)*1e-3); printf("\n"); }
        for (int oi = 0; oi < 14; ++oi) {
            const int sl = order[oi];
            double mn = 1e30, mxv = 0, av = 0; int cnt = 0;
            for (int b = 0; b < sms; ++b) { if (tl[b * 16 + sl] < t0) continue; const double v = (double)(tl[b * 16 + sl] - t0) * 1e-3; mn = fmin(mn, v); mxv = fmax(mxv, v); av += v; ++cnt; }
            printf("  timeline %-22s min %7.2f  avg %7.2f  max %7.2f us  (%d CTAs)\n", nm2[sl], mn, av / (cnt ? cnt : 1), mxv, cnt);
        }
    }
    {
        void* dws; const int64_t dwb = eco_dice_ws_bytes(3, 0); CK(cudaMalloc(&dws, dwb)); CK(cudaMemset(dws, 0, dwb));
        int64_t* cnt; double* soft; CK(cudaMalloc(&cnt, 64 * 8)); CK(cudaMalloc(&soft, 64 * 8));
        float t_d = time_us([&](int i) { eco_dice_counts(&vz[i % NSETS], &vg[i % NSETS], N, 3, HW, nullptr, 0, 0, dws, dwb, cnt, soft, 0, nullptr); }, iters);
        printf("dice_counts (LDG streaming of the same 8 B/element): %.2f us = %.0f GB/s\n", t_d, 8.0 * E / t_d * 1e-3);
    }
    printf("shipped: stats %.2f us  grad %.2f us  fused %.2f us\n", t_st, t_old, t_fu);
    printf("v2     : grad (no sums) %.2f us  fused %.2f us  -> %.1f GB/s (12 B/elem)\n", t_g2, t_f2, 12.0 * E / t_f2 * 1e-3);
    return 0;
}
