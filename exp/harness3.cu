// Experiment harness for the third-generation fused composite kernel (not part of the library): launches
// composite3_fused_v3_kernel directly, prints its in-kernel %globaltimer timeline (ECO_V2_TIMELINE) and times it.
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a [-DECO_V3_EXP_...] exp/harness3.cu \
//        ecologysemanticsegmentation_b200/csrc/build/eco_api.o -o exp/harness3
#define ECO_V2_TIMELINE 1
#include "../ecologysemanticsegmentation_b200/csrc/eco_composite.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline double urand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (rng_state >> 11) * (1.0 / 9007199254740992.0); }
static inline float nrand() { double u = urand() + 1e-12, v = urand(); return (float)(sqrt(-2 * log(u)) * cos(6.283185307179586 * v)); }

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
    float *z[NSETS], *g[NSETS], *o;
    for (int k = 0; k < NSETS; ++k) {
        CK(cudaMalloc(&z[k], E * 4)); CK(cudaMalloc(&g[k], E * 4));
        CK(cudaMemcpy(z[k], hz.data(), E * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(g[k], hg.data(), E * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&o, E * 4));
    double* scale_dev; float *losses, *up;
    CK(cudaMalloc(&scale_dev, 21 * 8)); CK(cudaMalloc(&losses, 7 * 4)); CK(cudaMalloc(&up, 7 * 4));
    const double r[3] = {1., 0.43197708, 0.22319692};
    double sc[21] = {2, 2, 2};
    { int t = 3; for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) { const double wi = 1 / r[i], wj = 1 / r[j], wd = 1 / (r[i] - r[j]);
        sc[t++] = 2 * wj; sc[t++] = 2 * wi; sc[t++] = 2 * wd; sc[t++] = 2 * wi; sc[t++] = 2 * wd; sc[t++] = 2 * wi * wi * wj; } }
    CK(cudaMemcpy(scale_dev, sc, sizeof(sc), cudaMemcpyHostToDevice));
    const float hup[7] = {0, 1, 0, 0, 1, 1, 1};
    CK(cudaMemcpy(up, hup, sizeof(hup), cudaMemcpyHostToDevice));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    using namespace eco::v2;
    CK(cudaFuncSetAttribute(composite3_fused_v3_kernel<float, float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Stage3<float, float>::kSmem));
    V3Ws* ws3; CK(cudaMalloc(&ws3, sizeof(V3Ws))); CK(cudaMemset(ws3, 0, sizeof(V3Ws)));
    unsigned int* status; CK(cudaMalloc(&status, 4)); CK(cudaMemset(status, 0, 4));
    auto launch = [&](int k) {
        EcoView vz{}, vg{}; vz.ptr = z[k]; vz.sn = C * HW; vz.sc = HW; vg = vz; vg.ptr = g[k];
        CompGradArgs ga{}; fill_comp(ga.a, &vz, &vg, N, HW, 4);
        ga.gx = o; ga.gx_sn = C * HW; ga.gx_sc = HW;
        XchArgs xch{}; xch.world = 1; xch.status = status;
        const double* sd = scale_dev; const float* u = up; unsigned int flags = 0;
        const float* prev = nullptr;
        void* args[] = {&ga, (void*)&sd, (void*)&u, &ws3, &losses, &flags, &xch, (void*)&prev};
        if (argc > 2) composite3_fused_v3_kernel<float, float, false><<<sms, kThreads3, Stage3<float, float>::kSmem>>>(ga, sd, u, ws3, losses, flags, xch, nullptr);
        else CK(cudaLaunchCooperativeKernel((const void*)composite3_fused_v3_kernel<float, float, false>, dim3(sms), dim3(kThreads3), args, Stage3<float, float>::kSmem, nullptr));
    };
    for (int i = 0; i < 5; ++i) launch(i % NSETS);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch(i % NSETS);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    float hl[7]; CK(cudaMemcpy(hl, losses, 28, cudaMemcpyDeviceToHost));
    printf("S=%d fused v3: %.2f us/step   losses:", S, ms * 1e3f / iters); for (int k = 0; k < 7; ++k) printf(" %.5f", hl[k]); printf("\n");
    { static unsigned long long z16[1024 * 16]; CK(cudaMemcpyToSymbol(g_timeline, z16, sizeof(z16))); }
    launch(0); CK(cudaDeviceSynchronize());
    static unsigned long long tl[1024 * 16];
    CK(cudaMemcpyFromSymbol(tl, g_timeline, sizeof(tl)));
    unsigned long long t0 = ~0ull; for (int b = 0; b < sms; ++b) t0 = tl[b * 16] < t0 ? tl[b * 16] : t0;
    const int order[15] = {0, 8, 1, 14, 15, 2, 7, 3, 13, 12, 4, 5, 9, 10, 6};
    const char* nm[16] = {"start", "pass1 loop end", "stats_finish end", "sums received", "coef ready", "pass2 loop end", "cta0 end", "extra lin tiles done", "lin start", "lin loop end", "lin arrived", "", "closed forms done", "layout done", "flush done", "sums in L2"};
    for (int oi = 0; oi < 15; ++oi) {
        const int sl = order[oi];
        double mn = 1e30, mxv = 0, av = 0; int cnt = 0;
        for (int b = 0; b < sms; ++b) { if (tl[b * 16 + sl] < t0) continue; const double v = (double)(tl[b * 16 + sl] - t0) * 1e-3; mn = fmin(mn, v); mxv = fmax(mxv, v); av += v; ++cnt; }
        printf("  timeline %-22s min %7.2f  avg %7.2f  max %7.2f us  (%d CTAs)\n", nm[sl], mn, av / (cnt ? cnt : 1), mxv, cnt);
    }
    return 0;
}
