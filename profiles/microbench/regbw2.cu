// Follow-up to regbw.cu (B200, sm_100a): which FFMA / FFMA2 operand shapes run at full FMA-pipe rate, and how
// scalar FFMA, FFMA2 and MUFU share issue slots.  Rates are lanes (fp32 FMAs) per clock per SM at the nominal
// max clock.  Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a regbw2.cu -o regbw2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int N = 12;

#define INIT_F(v) for (int i = 0; i < N; ++i) v[i] = 1.0f + 1e-3f * (threadIdx.x + i)
#define INIT_F2(v) for (int i = 0; i < N; ++i) v[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f + 2e-3f * i)
#define SINK_F(v) { float s = 0; for (int i = 0; i < N; ++i) s += v[i]; if (s == 12345.f) out[0] = s; }
#define SINK_F2(v) { float s = 0; for (int i = 0; i < N; ++i) s += v[i].x + v[i].y; if (s == 12345.f) out[0] = s; }

// MODE 0: acc += u*w (3 distinct changing)   1: acc += u*u (2 distinct)   2: acc = acc*u + acc (2 distinct)
// 3: acc += u*G, same G for the whole unrolled group (reuse cache)   4: horner r = r*u + K (K invariant register)
template <int MODE>
__global__ void __launch_bounds__(256) k_scalar(float* out, float a, float b) {
    float v[N], u[N];
    INIT_F(v); for (int i = 0; i < N; ++i) u[i] = 0.5f + 1e-4f * (threadIdx.x ^ i);
    float G = a;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) v[i] = fmaf(u[(i + 5) % N], u[(i + 7) % N], v[i]);
            if (MODE == 1) v[i] = fmaf(u[i], u[i], v[i]);
            if (MODE == 2) v[i] = fmaf(v[i], u[i], v[i]);
            if (MODE == 3) v[i] = fmaf(u[i], G, v[i]);
            if (MODE == 4) v[i] = fmaf(v[i], u[i], b);
        }
        G += b;
#pragma unroll
        for (int i = 0; i < N; i += 4) u[i] += b;  // keep u "changing" cheaply
    }
    SINK_F(v); SINK_F(u);
}
template <int MODE>
__global__ void __launch_bounds__(256) k_packed(float* out, float a, float b) {
    float2 v[N], u[N];
    INIT_F2(v); for (int i = 0; i < N; ++i) u[i] = make_float2(0.5f + 1e-4f * (threadIdx.x ^ i), 0.25f);
    float2 G = make_float2(a, a);
    const float2 b2 = make_float2(b, b);
    float sc[N];
    for (int i = 0; i < N; ++i) sc[i] = 0.999f + 1e-5f * (threadIdx.x + 3 * i) * a;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE >= 7) {
#pragma unroll
            for (int i = 0; i < N; i += 4) sc[i] += b * 1e-6f;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) v[i] = __ffma2_rn(u[(i + 5) % N], u[(i + 7) % N], v[i]);
            if (MODE == 1) v[i] = __ffma2_rn(u[i], u[i], v[i]);
            if (MODE == 2) v[i] = __ffma2_rn(v[i], u[i], v[i]);
            if (MODE == 3) v[i] = __ffma2_rn(u[i], G, v[i]);
            if (MODE == 4) v[i] = __ffma2_rn(v[i], u[i], b2);
            if (MODE == 5) v[i] = __fmul2_rn(v[i], u[i]);
            if (MODE == 6) v[i] = __fadd2_rn(v[i], u[i]);
            if (MODE == 7) v[i] = __ffma2_rn(v[i], u[i], make_float2(sc[i], sc[i]));            // vec, vec, scalar
            if (MODE == 8) v[i] = __ffma2_rn(v[i], make_float2(sc[i], sc[i]), make_float2(sc[(i + 3) % N], sc[(i + 3) % N]));  // vec, scalar, scalar
            if (MODE == 9) v[i] = __ffma2_rn(u[i], make_float2(sc[i], sc[i]), v[i]);            // vec, scalar, vec(acc)
            if (MODE == 10) v[i] = __ffma2_rn(v[i], make_float2(sc[i], sc[i]), make_float2(0.25f, 0.25f));  // vec, scalar, imm
        }
        G = __fadd2_rn(G, b2);
#pragma unroll
        for (int i = 0; i < N; i += 4) u[i] = __fadd2_rn(u[i], b2);
    }
    SINK_F2(v); SINK_F2(u);
    if (MODE >= 7) { float t = 0; for (int i = 0; i < N; ++i) t += sc[i]; if (t == 12345.f) out[1] = t; }
}
// interleave: per group of 3 instructions, P packed 3-distinct FFMA2 and S scalar 3-distinct FFMA
template <int P, int S>
__global__ void __launch_bounds__(256) k_mixps(float* out, float a, float b) {
    float2 v[N], u[N];
    float w[N], x[N];
    INIT_F2(v); INIT_F(w);
    for (int i = 0; i < N; ++i) { u[i] = make_float2(0.5f + 1e-4f * (threadIdx.x ^ i), 0.25f); x[i] = 0.5f + 1e-4f * i; }
    const float2 b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int p = 0; p < P; ++p) v[(i + p) % N] = __ffma2_rn(u[(i + 5 + p) % N], u[(i + 7 + p) % N], v[(i + p) % N]);
#pragma unroll
            for (int s = 0; s < S; ++s) w[(i + s) % N] = fmaf(x[(i + 5 + s) % N], x[(i + 7 + s) % N], w[(i + s) % N]);
        }
#pragma unroll
        for (int i = 0; i < N; i += 4) { u[i] = __fadd2_rn(u[i], b2); x[i] += b; }
    }
    SINK_F2(v); SINK_F2(u); SINK_F(w); SINK_F(x);
}
// FFMA2 (3 distinct) with M MUFU per N packed FMAs
template <int M, bool PACKED>
__global__ void __launch_bounds__(256) k_mufumix(float* out, float a, float b) {
    float2 v[N], u[N];
    float m[8];
    INIT_F2(v);
    for (int i = 0; i < N; ++i) u[i] = make_float2(0.5f + 1e-4f * (threadIdx.x ^ i), 0.25f);
    for (int i = 0; i < 8; ++i) m[i] = 1.0f + 0.001f * (threadIdx.x + i);
    const float2 b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (PACKED) v[i] = __ffma2_rn(u[(i + 5) % N], u[(i + 7) % N], v[i]);
            else { v[i].x = fmaf(u[(i + 5) % N].x, u[(i + 7) % N].x, v[i].x); v[i].y = fmaf(u[(i + 5) % N].y, u[(i + 7) % N].y, v[i].y); }
            if (i < M) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(m[i % 8]));
        }
#pragma unroll
        for (int i = 0; i < N; i += 4) u[i] = __fadd2_rn(u[i], b2);
    }
    SINK_F2(v); SINK_F2(u);
    float s = 0; for (int i = 0; i < 8; ++i) s += m[i]; if (s == 12345.f) out[0] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float* out; cudaMalloc(&out, 8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;
    const double warps = (double)blocks * 8;
    // warp-instr per clk per SM for `per_iter` instructions per loop iteration
    auto rate = [&](float ms, double per_iter) { return warps * (double)ITERS * per_iter / (ms * 1e-3) / ((double)clk * 1e3) / sms; };
    float ms;
    const char* names[] = {"acc+=u*w 3 distinct", "acc+=u*u 2 distinct", "acc=acc*u+acc 2 distinct", "acc+=u*G shared G", "r=r*u+K horner"};
#define RUN_S(M) ms = timeit([&] { k_scalar<M><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("FFMA  %-26s: %.2f instr/clk/SM (%.0f lanes)\n", names[M], rate(ms, N), 32 * rate(ms, N));
    RUN_S(0) RUN_S(1) RUN_S(2) RUN_S(3) RUN_S(4)
#define RUN_P(M, nm) ms = timeit([&] { k_packed<M><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("%-32s: %.2f instr/clk/SM (%.0f lanes)\n", nm, rate(ms, N), 64 * rate(ms, N));
    RUN_P(0, "FFMA2 acc+=u*w 3 distinct") RUN_P(1, "FFMA2 acc+=u*u 2 distinct") RUN_P(2, "FFMA2 acc=acc*u+acc 2 distinct")
    RUN_P(3, "FFMA2 acc+=u*G shared G") RUN_P(4, "FFMA2 r=r*u+K horner") RUN_P(5, "FMUL2 v*=u") RUN_P(6, "FADD2 v+=u")
    RUN_P(7, "FFMA2 vec,vec,scalar") RUN_P(8, "FFMA2 vec,scalar,scalar") RUN_P(9, "FFMA2 vec,scalar,acc") RUN_P(10, "FFMA2 vec,scalar,imm")
#define RUN_M(P, S) ms = timeit([&] { k_mixps<P, S><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("mix %d FFMA2 + %d FFMA (3 distinct): %.2f instr/clk/SM, %.0f lanes\n", P, S, rate(ms, N * (P + S)), 32 * rate(ms, N * (2 * P + S)));
    RUN_M(1, 1) RUN_M(1, 2) RUN_M(2, 1) RUN_M(1, 0) RUN_M(0, 1)
#define RUN_U(M, PK) ms = timeit([&] { k_mufumix<M, PK><<<blocks, 256>>>(out, 1.0001f, 0.5f); }); printf("%s x12 (3 distinct) + %d MUFU: %.2f instr/clk/SM, %.0f fma lanes, %.1f mufu lanes\n", PK ? "FFMA2" : "2xFFMA", M, rate(ms, (PK ? N : 2 * N) + M), 32 * rate(ms, 2 * N), 32 * rate(ms, M));
    RUN_U(0, true) RUN_U(2, true) RUN_U(3, true) RUN_U(4, true) RUN_U(6, true)
    RUN_U(0, false) RUN_U(2, false) RUN_U(3, false) RUN_U(4, false)
    return 0;
}
