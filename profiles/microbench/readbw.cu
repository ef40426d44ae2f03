// Read-bandwidth floor for a ONE-SHOT sweep of a small working set (28 / 85 / 340 / 1360 MB) on B200: what can pass 1 of
// the fused loss kernels hope for at cfg2's size?  Plain 128-bit streaming loads (ld.global.nc.L1::no_allocate), U loads
// in flight per thread, persistent grid of 148 x B CTAs x 256 threads, buffers rotated so nothing is L2-resident.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/microbench/readbw.cu -o profiles/microbench/readbw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <int U>
__global__ void __launch_bounds__(256) sweep(const float4* __restrict__ src, size_t n4, float* __restrict__ sink) {
    // every CTA owns one contiguous range; a thread keeps U independent 16-byte loads in flight
    const size_t per_cta = (n4 + gridDim.x - 1) / gridDim.x;
    size_t lo = per_cta * blockIdx.x, hi = lo + per_cta;
    if (hi > n4) hi = n4;
    float acc = 0.f;
    for (size_t i = lo + threadIdx.x; i < hi; i += (size_t)256 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t j = i + (size_t)u * 256;
            v[u] = j < hi ? ldg_stream(src + j) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) *sink = acc;
}

template <int U>
static void run(const char* name, float4** bufs, int nbuf, size_t bytes, int ctas_per_sm, float* sink) {
    const size_t n4 = bytes / 16;
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) sweep<U><<<grid, 256>>>(bufs[i % nbuf], n4, sink);
    CK(cudaDeviceSynchronize());
    const int iters = 100;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) sweep<U><<<grid, 256>>>(bufs[i % nbuf], n4, sink);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / iters;
    printf("%-8s %7.1f MB  U=%d  %d CTAs/SM : %7.2f us  %7.1f GB/s\n", name, bytes / 1e6, U, ctas_per_sm, us, bytes / us / 1e3);
}

int main() {
    float* sink; CK(cudaMalloc(&sink, 4));
    const size_t sizes[] = {28311552ull, 84934656ull, 339738624ull, 1358954496ull};
    const char* names[] = {"cfg1", "cfg2", "cfg4shard", "cfg3"};
    for (int s = 0; s < 4; ++s) {
        const size_t bytes = sizes[s];
        int nbuf = (int)(600000000ull / bytes) + 1;   // > 4x L2 in rotation
        if (nbuf < 2) nbuf = 2;
        if (nbuf > 24) nbuf = 24;
        float4* bufs[24];
        for (int k = 0; k < nbuf; ++k) { CK(cudaMalloc(&bufs[k], bytes)); CK(cudaMemset(bufs[k], 0, bytes)); }
        run<4>(names[s], bufs, nbuf, bytes, 4, sink);
        run<4>(names[s], bufs, nbuf, bytes, 8, sink);
        run<8>(names[s], bufs, nbuf, bytes, 4, sink);
        run<8>(names[s], bufs, nbuf, bytes, 8, sink);
        for (int k = 0; k < nbuf; ++k) CK(cudaFree(bufs[k]));
    }
    return 0;
}
