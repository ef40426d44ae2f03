// Prediction / label tensors -> byte masks for the result dumps that follow the scoring:
//   ess/test_multiclass.py:58      out = F.sigmoid(net(x))
//   ess/test_multiclass.py:68-69   out[out > T] = 1 ; out[out != 1] = 0            (optional threshold rule)
//   ess/test_multiclass.py:90-92   (t.numpy() * 255).astype(np.uint8)              (images / labels / outputs)
//   ess/test_video.py:129-130      (output_image * 255).astype(np.uint8)
// The reference moves fp32 tensors to the host and converts there (3 sweeps + a 4 B/element copy); here sigmoid,
// threshold, the fp32 multiply by 255 and the truncation run in ONE pass that reads 4 B and writes 1 B per element,
// so the device->host copy of a dump shrinks 4x.  HBM-bound: 5 B/element; one thread turns 16 consecutive elements
// (four 128-bit loads in flight) into one 128-bit store.
#include "eco_common.cuh"

namespace eco {

struct MaskArgs {
    const void* x;
    int64_t sn, sc;
    int32_t N, C;
    int64_t HW;
    float thr;
    int32_t use_thr;
    int32_t probs;
    uint8_t* out;
};

// one element: bit-compatible with ATen's CUDA sigmoid, numpy's float32 `* 255` and the truncating uint8 cast
__device__ __forceinline__ uint32_t mask_byte(float v, const MaskArgs& p) {
    float q = p.probs ? v : sigmoid_exact(v);
    if (p.use_thr) q = (q > p.thr || q == 1.0f) ? 1.0f : 0.0f;   // `> T` -> 1, then everything that is not exactly 1 -> 0
    return (uint32_t)(int)__fmul_rn(q, 255.0f) & 0xffu;
}

template <typename T, int VEC /* 16 or 1 */>
__global__ void __launch_bounds__(256) masks_u8_kernel(MaskArgs p) {
    const int64_t per_plane = p.HW / VEC;
    const int64_t total = (int64_t)p.N * p.C * per_plane;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t plane = i / per_plane, e = (i - plane * per_plane) * VEC;
        const int64_t n = plane / p.C, c = plane - n * p.C;
        const T* src = reinterpret_cast<const T*>(p.x) + n * p.sn + c * p.sc + e;
        uint8_t* dst = p.out + plane * p.HW + e;
        if constexpr (VEC == 16) {
            float v[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) Vec4<T>::load(src + 4 * k, v[k]);
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = mask_byte(v[k][0], p) | (mask_byte(v[k][1], p) << 8) | (mask_byte(v[k][2], p) << 16) |
                       (mask_byte(v[k][3], p) << 24);
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                         "r"(w[3])
                         : "memory");
        } else {
            *dst = (uint8_t)mask_byte(Vec4<T>::load1(src), p);
        }
    }
}

}  // namespace eco

using namespace eco;

extern "C" int eco_masks_u8(const EcoView* x, int32_t N, int32_t C, int64_t HW, float threshold, int32_t use_threshold,
                            int32_t x_is_prob, uint8_t* out, int device, void* stream) {
    if (N <= 0 || C <= 0 || HW <= 0) { set_error("empty input (N=%d C=%d HW=%lld)", N, C, (long long)HW); return -2; }
    if (!x || !x->ptr || !out) { set_error("null input view or output"); return -1; }
    if (x->dtype != ECO_F32 && x->dtype != ECO_BF16) { set_error("eco_masks_u8: input must be f32 or bf16"); return -4; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    MaskArgs p{x->ptr, x->sn, x->sc, N, C, HW, threshold, use_threshold, x_is_prob, out};
    const int64_t esz = x->dtype == ECO_BF16 ? 2 : 4;
    const bool v16 = HW % 16 == 0 && x->sn % 4 == 0 && x->sc % 4 == 0 && reinterpret_cast<uintptr_t>(x->ptr) % (4 * esz) == 0 &&
                     reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -10;
    const int64_t total = (int64_t)N * C * (HW / (v16 ? 16 : 1));
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;   // resident wave, grid-stride beyond it
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x->dtype == ECO_F32) {
        if (v16) masks_u8_kernel<float, 16><<<(unsigned)grid, 256, 0, st>>>(p);
        else masks_u8_kernel<float, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    } else {
        if (v16) masks_u8_kernel<__nv_bfloat16, 16><<<(unsigned)grid, 256, 0, st>>>(p);
        else masks_u8_kernel<__nv_bfloat16, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    }
    return check_cuda(cudaGetLastError(), "masks_u8_kernel launch");
}
