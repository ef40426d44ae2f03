// Does FFMA / FFMA2 throughput depend on how many DISTINCT register operands an instruction reads?
// (B200, sm_100a).  Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a regbw.cu -o regbw
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int N = 12;

// MODE 0: v = v*a + b (a, b loop-invariant: 1 changing operand)   MODE 1: v[i] = v[i]*v[j] + v[k] (3 distinct, all changing)
template <int MODE>
__global__ void __launch_bounds__(256) k_scalar(float* out, float a, float b) {
    float v[N];
    for (int i = 0; i < N; ++i) v[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) v[i] = fmaf(v[i], a, b);
            else v[i] = fmaf(v[(i + 5) % N], v[(i + 7) % N], v[i]);
        }
    float s = 0; for (int i = 0; i < N; ++i) s += v[i];
    if (s == 12345.f) out[0] = s;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_packed(float* out, float a, float b) {
    float2 v[N];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < N; ++i) v[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f + 2e-3f * i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) v[i] = __ffma2_rn(v[i], a2, b2);
            else if (MODE == 1) v[i] = __ffma2_rn(v[(i + 5) % N], v[(i + 7) % N], v[i]);
            else v[i] = __ffma2_rn(v[(i + 5) % N], a2, v[i]);   // 2 distinct changing + 1 invariant
        }
    float s = 0; for (int i = 0; i < N; ++i) s += v[i].x + v[i].y;
    if (s == 12345.f) out[0] = s;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_packed_add(float* out, float a) {
    float2 v[N];
    const float2 a2 = make_float2(a, a);
    for (int i = 0; i < N; ++i) v[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f + 2e-3f * i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) v[i] = __fadd2_rn(v[i], a2);
            else v[i] = __fadd2_rn(v[(i + 5) % N], v[i]);
        }
    float s = 0; for (int i = 0; i < N; ++i) s += v[i].x + v[i].y;
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
    const int blocks = sms * 8;
    const double warps = (double)blocks * 8;
    auto rate = [&](float ms) { return warps * (double)ITERS * N / (ms * 1e-3) / ((double)clk * 1e3) / sms; };
    float ms;
    ms = timeit([&] { k_scalar<0><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA  1 changing operand : %.2f warp-instr/clk/SM\n", rate(ms));
    ms = timeit([&] { k_scalar<1><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA  3 distinct operands: %.2f warp-instr/clk/SM\n", rate(ms));
    ms = timeit([&] { k_packed<0><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA2 1 changing operand : %.2f warp-instr/clk/SM (%.0f fma lanes/clk/SM)\n", rate(ms), 64 * rate(ms));
    ms = timeit([&] { k_packed<2><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA2 2 changing operands: %.2f warp-instr/clk/SM (%.0f fma lanes/clk/SM)\n", rate(ms), 64 * rate(ms));
    ms = timeit([&] { k_packed<1><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA2 3 distinct operands: %.2f warp-instr/clk/SM (%.0f fma lanes/clk/SM)\n", rate(ms), 64 * rate(ms));
    ms = timeit([&] { k_packed_add<0><<<blocks, 256>>>(out, 0.5f); }); printf("FADD2 1 changing operand : %.2f warp-instr/clk/SM\n", rate(ms));
    ms = timeit([&] { k_packed_add<1><<<blocks, 256>>>(out, 0.5f); }); printf("FADD2 2 distinct operands: %.2f warp-instr/clk/SM\n", rate(ms));
    return 0;
}
