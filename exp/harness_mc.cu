// Experiment harness for the one-launch plain 3-organ step (not part of the library): launches
// multiclass3_fused_v2_kernel directly, prints its in-kernel %globaltimer timeline (ECO_V2_TIMELINE) and times it.
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a exp/harness_mc.cu \
//        ecologysemanticsegmentation_b200/csrc/build/eco_api.o -o exp/harness_mc
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
    for (int64_t i = 0; i < E; ++i) { hz[i] = nrand(); hg[i] = urand() < 0.4; }
    float *z[NSETS], *g[NSETS], *o[NSETS];
    for (int k = 0; k < NSETS; ++k) {
        CK(cudaMalloc(&z[k], E * 4)); CK(cudaMalloc(&g[k], E * 4)); CK(cudaMalloc(&o[k], E * 4));
        CK(cudaMemcpy(z[k], hz.data(), E * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(g[k], hg.data(), E * 4, cudaMemcpyHostToDevice));
    }
    float *losses, *up;
    CK(cudaMalloc(&losses, 7 * 4)); CK(cudaMalloc(&up, 7 * 4));
    const float hup[7] = {0, 1, 0, 0, 1, 1, 1};
    CK(cudaMemcpy(up, hup, sizeof(hup), cudaMemcpyHostToDevice));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    using namespace eco::v2;
    CK(cudaFuncSetAttribute(multiclass3_fused_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    V2Ws* ws2; CK(cudaMalloc(&ws2, sizeof(V2Ws))); CK(cudaMemset(ws2, 0, sizeof(V2Ws)));
    auto launch = [&](int k) {
        EcoView vz{}, vg{}; vz.ptr = z[k]; vz.sn = C * HW; vz.sc = HW; vg = vz; vg.ptr = g[k];
        CompGradArgs ga{}; fill_comp(ga.a, &vz, &vg, N, HW, 4);
        ga.gx = o[k]; ga.gx_sn = C * HW; ga.gx_sc = HW;
        double scale = 1.0; const float* u = up;
        void* args[] = {&ga, &scale, (void*)&u, &ws2, &losses};
        CK(cudaLaunchCooperativeKernel((const void*)multiclass3_fused_v2_kernel, dim3(sms), dim3(kThreads), args, kSmemBytes, nullptr));
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
    printf("S=%d multiclass step: %.2f us/step   losses:", S, ms * 1e3f / iters); for (int k = 0; k < 7; ++k) printf(" %.5f", hl[k]); printf("\n");
    { static unsigned long long z16[1024 * 16]; CK(cudaMemcpyToSymbol(g_timeline, z16, sizeof(z16))); }
    launch(1); CK(cudaDeviceSynchronize());
    static unsigned long long tl[1024 * 16];
    CK(cudaMemcpyFromSymbol(tl, g_timeline, sizeof(tl)));
    unsigned long long t0 = ~0ull; for (int b = 0; b < sms; ++b) t0 = tl[b * 16] < t0 ? tl[b * 16] : t0;
    const char* nm[6] = {"start", "pass1 loop end", "sums in L2", "all arrived", "coef ready", "pass2 loop end"};
    for (int sl = 0; sl < 6; ++sl) {
        double mn = 1e30, mxv = 0, av = 0; int cnt = 0;
        for (int b = 0; b < sms; ++b) { if (tl[b * 16 + sl] < t0) continue; const double v = (double)(tl[b * 16 + sl] - t0) * 1e-3; mn = fmin(mn, v); mxv = fmax(mxv, v); av += v; ++cnt; }
        printf("  timeline %-18s min %7.2f  avg %7.2f  max %7.2f us  (%d CTAs)\n", nm[sl], mn, av / (cnt ? cnt : 1), mxv, cnt);
    }
    return 0;
}
