// Label union / un-union step adjacent to the loss path, in place:
//   ess/utils/subsets_union.py:8-32  return_union_sets_descending_order(ann, exclude_indices, reverse)  (class dim)
//   ess/train_multiclass.py:32-45    its batch-dim twin, the one train() calls at :110
// The tensor is viewed as [outer][K][inner] with `inner` contiguous; one thread owns one (outer, inner) column
// of K values, walks it once from the back (the suffix sum / the already-updated neighbour is a running scalar),
// so every element is read and written exactly once: 8 B/element, HBM-bound.
#include "eco_common.cuh"

namespace eco {

struct UnionArgs {
    void* data;
    int64_t outer, inner, stride_outer, stride_k;
    int32_t K;
    uint64_t exclude_mask;
    int32_t reverse;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256) union_sets_kernel(UnionArgs p) {
    const int64_t cols = p.outer * (p.inner / VEC);
    for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = col / (p.inner / VEC), i = (col - o * (p.inner / VEC)) * VEC;
        T* base = reinterpret_cast<T*>(p.data) + o * p.stride_outer + i;
        float run[VEC];  // forward: suffix sum of the ORIGINAL values; reverse: the UPDATED value of k+1
#pragma unroll
        for (int v = 0; v < VEC; ++v) run[v] = 0.f;
        for (int k = p.K - 1; k >= 0; --k) {
            T* ptr = base + (int64_t)k * p.stride_k;
            float x[VEC];
            if (VEC == 4) Vec4<T>::load(ptr, reinterpret_cast<float(&)[4]>(x));
            else x[0] = Vec4<T>::load1(ptr);
            const bool touched = k < p.K - 1 && !((p.exclude_mask >> k) & 1ull);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (!p.reverse) {
                    run[v] += x[v];
                    float out = touched ? run[v] : x[v];
                    x[v] = out > 1.0f ? 1.0f : out;      // ann[ann > 1] = 1 (NaN stays NaN, like the reference)
                } else {
                    if (touched) x[v] = fabsf(x[v] - run[v]);
                    run[v] = x[v];
                }
            }
            if (VEC == 4) Vec4<T>::store(ptr, reinterpret_cast<float(&)[4]>(x));
            else Vec4<T>::store1(ptr, x[0]);
        }
    }
}

}  // namespace eco

using namespace eco;

extern "C" int eco_union_sets(void* data, int32_t dtype, int64_t outer, int32_t K, int64_t inner, int64_t stride_outer,
                              int64_t stride_k, uint64_t exclude_mask, int32_t reverse, int device, void* stream) {
    if (outer <= 0 || K <= 0 || inner <= 0) { set_error("empty input (outer=%lld K=%d inner=%lld)", (long long)outer, K, (long long)inner); return -2; }
    if (!data) { set_error("null data"); return -1; }
    if (K > 64) { set_error("eco_union_sets: K=%d exceeds the 64-entry exclude mask", K); return -3; }
    if (dtype != ECO_F32 && dtype != ECO_BF16) { set_error("unsupported dtype code"); return -4; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    UnionArgs p{data, outer, inner, stride_outer, stride_k, K, exclude_mask, reverse};
    const int64_t esz = dtype == ECO_BF16 ? 2 : 4;
    const bool v4 = inner % 4 == 0 && stride_outer % 4 == 0 && stride_k % 4 == 0 && reinterpret_cast<uintptr_t>(data) % (4 * esz) == 0;
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -10;
    const int64_t cols = outer * (inner / (v4 ? 4 : 1));
    int64_t grid = (cols + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ECO_F32) {
        if (v4) union_sets_kernel<float, 4><<<(unsigned)grid, 256, 0, st>>>(p);
        else union_sets_kernel<float, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    } else {
        if (v4) union_sets_kernel<__nv_bfloat16, 4><<<(unsigned)grid, 256, 0, st>>>(p);
        else union_sets_kernel<__nv_bfloat16, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    }
    return check_cuda(cudaGetLastError(), "union_sets_kernel launch");
}
