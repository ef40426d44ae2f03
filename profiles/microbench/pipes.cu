// Pipe-rate microbenchmark for the composite kernel design (B200, sm_100a):
//   scalar FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU (ex2/lg2/rcp/sqrt), and their overlap.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu -o pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int NCHAIN = 8;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = fmaf(v[i], a, b);
    float s = 0; for (int i = 0; i < NCHAIN; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
    float2 v[NCHAIN];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < NCHAIN; ++i) v[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
    float s = 0; for (int i = 0; i < NCHAIN; ++i) s += v[i].x + v[i].y;
    if (s == 12345.f) out[0] = s;
}
template <int OP>
__global__ void __launch_bounds__(256) k_mufu(float* out, float a) {
    float v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = 1.0f + 0.001f * (threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
        }
    float s = 0; for (int i = 0; i < NCHAIN; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}
// R packed FMAs per MUFU, independent chains: does the MUFU hide under the FMA stream?
template <int R, bool PACKED>
__global__ void __launch_bounds__(256) k_mix(float* out, float a, float b) {
    float m[4];
    float2 v[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < 4; ++i) m[i] = 1.0f + 0.001f * (threadIdx.x + i);
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (PACKED) v[(r * 4 + i) % 8] = __ffma2_rn(v[(r * 4 + i) % 8], a2, b2);
                else { v[(r * 4 + i) % 8].x = fmaf(v[(r * 4 + i) % 8].x, a, b); }
            }
    }
    float s = 0; for (int i = 0; i < 4; ++i) s += m[i]; for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 12345.f) out[0] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float* out; cudaMalloc(&out, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;  // 2048 threads / SM
    const double warps = (double)blocks * 8;
    auto rate = [&](float ms, double per_thread_ops) {  // warp-instr per clk per SM (at nominal max clock)
        return warps * per_thread_ops / (ms * 1e-3) / ((double)clk * 1e3) / sms;
    };
    printf("SMs %d, clock %d kHz, %d blocks x 256\n", sms, clk, blocks);
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, 256>>>(out, 1.0001f, 0.5f); });
    printf("FFMA   : %.3f ms  -> %.2f warp-instr/clk/SM (%.1f lanes/clk/SM)\n", ms, rate(ms, (double)ITERS * NCHAIN), 32 * rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_ffma2<<<blocks, 256>>>(out, 1.0001f, 0.5f); });
    printf("FFMA2  : %.3f ms  -> %.2f warp-instr/clk/SM (%.1f fma lanes/clk/SM)\n", ms, rate(ms, (double)ITERS * NCHAIN), 64 * rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_mufu<0><<<blocks, 256>>>(out, 1.f); });
    printf("MUFU.EX2 : %.3f ms -> %.3f warp-instr/clk/SM (%.1f lanes/clk/SM)\n", ms, rate(ms, (double)ITERS * NCHAIN), 32 * rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_mufu<1><<<blocks, 256>>>(out, 1.f); });
    printf("MUFU.LG2 : %.3f ms -> %.3f warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_mufu<2><<<blocks, 256>>>(out, 1.f); });
    printf("MUFU.RCP : %.3f ms -> %.3f warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_mufu<3><<<blocks, 256>>>(out, 1.f); });
    printf("MUFU.SQRT: %.3f ms -> %.3f warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * NCHAIN));
    ms = timeit([&] { k_mix<4, true><<<blocks, 256>>>(out, 1.0001f, 0.5f); });
    printf("mix 4 MUFU + 16 FFMA2 / iter: %.3f ms -> %.2f total warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * 20));
    ms = timeit([&] { k_mix<8, true><<<blocks, 256>>>(out, 1.0001f, 0.5f); });
    printf("mix 4 MUFU + 32 FFMA2 / iter: %.3f ms -> %.2f total warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * 36));
    ms = timeit([&] { k_mix<8, false><<<blocks, 256>>>(out, 1.0001f, 0.5f); });
    printf("mix 4 MUFU + 32 FFMA  / iter: %.3f ms -> %.2f total warp-instr/clk/SM\n", ms, rate(ms, (double)ITERS * 36));
    return 0;
}
