// Recorded experiment for VERDICT r1 item 2: could pass 1 of the composite kernel (59 sums per pixel) run as a Gram-matrix
// contraction F^T F on tcgen05?  Before any MMA is issued, the operands have to be in shared memory: per pixel a feature
// vector of >= 36 fp32 values (x_c, x_c^2, and per organ pair d, m1, q, m3, u1..u3, u1^2..u3^2) split into two TF32 halves
// (hi, lo) for 1e-5 accuracy = 72 values = 288 bytes, against 24 bytes of input.  This kernel measures ONLY that staging
// (the per-pixel sigmoids and feature products are left out, the shared-memory writes are what is timed), at cfg2's volume
// (54 x 256 x 256 pixels) with 148 CTAs x 512 threads, and compares it with the 20 us of the scalar pass 1.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/microbench/gram_staging.cu -o profiles/microbench/gram_staging
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kFeat = 72;                     // 36 features x (hi, lo)
constexpr int kThreads = 512;
constexpr int kTilePix = 512;                 // pixels staged per CTA iteration: 512 x 288 B = 144 KB

__global__ void __launch_bounds__(kThreads, 1) stage_kernel(const float* __restrict__ z, int64_t npix, float* __restrict__ sink) {
    extern __shared__ float4 tile[];          // [kFeat / 4][kTilePix] float4: a K-major operand tile
    float acc = 0.f;
    for (int64_t p0 = (int64_t)blockIdx.x * kTilePix; p0 < npix; p0 += (int64_t)gridDim.x * kTilePix) {
        const int64_t p = p0 + threadIdx.x;
        const float x = p < npix ? z[p] : 0.f;
#pragma unroll
        for (int f = 0; f < kFeat / 4; ++f) {
            // stand-in for four (hi, lo) feature halves: two FMA-class instructions per value, as a real split costs
            const float a = fmaf(x, 0.5f + f, 1.0f), b = a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);
            tile[f * kTilePix + threadIdx.x] = make_float4(a, b, fmaf(a, x, b), fmaf(b, x, a));
        }
        __syncthreads();                      // here the MMAs would be issued and waited for
        acc += reinterpret_cast<const float*>(tile)[(threadIdx.x * 37) % (kFeat * kTilePix)];
        __syncthreads();
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const int64_t npix = 54ll * 256 * 256;
    float *z, *sink;
    cudaMalloc(&z, npix * 4); cudaMalloc(&sink, 4);
    cudaMemset(z, 0, npix * 4);
    const int smem = kFeat * kTilePix * 4;
    cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) stage_kernel<<<148, kThreads, smem>>>(z, npix, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 50; ++i) stage_kernel<<<148, kThreads, smem>>>(z, npix, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("staging 72 fp32 values (288 B) per pixel for %lld pixels: %.2f us per sweep (%s)\n", (long long)npix, ms * 1e3f / 50,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
