// Prediction / label tensors -> byte masks for the result dumps that follow the scoring:
//   ess/test_multiclass.py:58      out = F.sigmoid(net(x))
//   ess/test_multiclass.py:68-69   out[out > T] = 1 ; out[out != 1] = 0            (optional threshold rule)
//   ess/test_multiclass.py:90-92   (t.numpy() * 255).astype(np.uint8)              (images / labels / outputs)
//   ess/test_video.py:129-130      (output_image * 255).astype(np.uint8)
// The reference moves fp32 tensors to the host and converts there (3 sweeps + a 4 B/element copy); here sigmoid,
// threshold, the fp32 multiply by 255 and the truncation run in ONE pass that reads 4 B and writes 1 B per element,
// so the device->host copy of a dump shrinks 4x.  HBM-bound: 5 B/element.
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

// One element, bit-compatible with ATen's CUDA sigmoid, numpy's float32 `* 255` and the truncating uint8 cast.
//  * The sigmoid is the 4-instruction MUFU form (|error| < 2.5e-7); ATen's exact bits (expf + IEEE division) are
//    recomputed only where that could change the byte: within kNearThr of the threshold, or with q * 255 within
//    kNearInt of an integer (~0.06 % of the elements).  A thread first forms all 16 bytes branch-free (the MUFU
//    latencies of the 16 elements overlap) while collecting a bit mask of the doubtful ones, then redoes those.
//  * No F2I / FRND (they share the 16-lane XU pipe with the two MUFU of the sigmoid and would set the pace): for
//    0 <= s < 2^23 the round-down add  m = s + 2^23  holds floor(s) in its low mantissa bits, and m - (2^23 - 0.5) is
//    floor(s) + 0.5 exactly, which gives the distance of s to the nearest integer with two more FADDs.
constexpr float kNearThr = 4e-6f;
constexpr float kNearInt = 3e-4f;   // 255 * (fast vs exact sigmoid, < 3.3e-7) = 8.4e-5: 3.5x margin
constexpr float kTwo23 = 8388608.0f;
enum : int { kModeProb = 0, kModeProbThr = 1, kModeSig = 2, kModeSigThr = 3 };

__device__ __forceinline__ uint32_t floor_byte(float s_nonneg, float& m) {
    m = __fadd_rd(s_nonneg, kTwo23);
    return __float_as_uint(m) & 0xffu;
}
// `> T` -> 1, then everything that is not exactly 1 -> 0; times 255
__device__ __forceinline__ uint32_t thr_byte(float q, float thr) { return (q > thr || q == 1.0f) ? 255u : 0u; }

template <int MODE>
__device__ __forceinline__ uint32_t fast_byte(float v, const MaskArgs& p, bool& doubtful) {
    doubtful = false;
    if (MODE == kModeProbThr) return thr_byte(v, p.thr);
    if (MODE == kModeProb) {
        // probabilities / labels / images as they are: trunc(fp32(v * 255)) toward zero, low 8 bits (what the x86 cast does)
        const float s = __fmul_rn(v, 255.0f);
        float m;
        const uint32_t k = floor_byte(fabsf(s), m);
        return s < 0.f ? ((0u - k) & 0xffu) : k;
    }
    const float q = sigmoid_fast(v);
    if (MODE == kModeSigThr) {
        doubtful = fabsf(q - p.thr) < kNearThr || q > 0.99999f;
        return q > p.thr ? 255u : 0u;
    }
    float m;
    const float s = __fmul_rn(q, 255.0f);
    const uint32_t k = floor_byte(s, m);
    doubtful = fabsf(s - (m - (kTwo23 - 0.5f))) > 0.5f - kNearInt;   // s within kNearInt of an integer
    return k;
}
template <int MODE>
__device__ __noinline__ uint32_t exact_byte(float v, float thr) {
    const float q = sigmoid_exact(v);
    if (MODE == kModeSigThr) return thr_byte(q, thr);
    float m;
    return floor_byte(__fmul_rn(q, 255.0f), m);
}

// VEC == 16: a warp owns 512 consecutive elements of a plane; load k of a lane covers elements 128 k + 4 lane .. + 3
// (every load instruction is one fully coalesced 512-byte request, four in flight per thread) and its four bytes go
// out as one 32-bit store (128 coalesced bytes per warp-store).
template <typename T, int VEC /* 16 or 1 */, int MODE>
__global__ void __launch_bounds__(256) masks_u8_kernel(MaskArgs p) {
    if constexpr (VEC == 16) {
        // every warp walks ONE contiguous range of 512-element chunks and tracks (image, channel, chunk) incrementally:
        // no per-iteration 64-bit division (it would cost more instructions than the 16 elements themselves)
        const int lane = threadIdx.x & 31;
        const int64_t cpp = p.HW / 512;                                  // chunks per plane
        const int64_t total = (int64_t)p.N * p.C * cpp;
        const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
        const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        int64_t i = total * w / warps;
        const int64_t i_end = total * (w + 1) / warps;
        if (i >= i_end) return;
        int64_t plane = i / cpp;
        int64_t chunk = i - plane * cpp;
        int64_t n = plane / p.C;
        int c = (int)(plane - n * p.C);
        const T* src = reinterpret_cast<const T*>(p.x) + n * p.sn + c * p.sc + chunk * 512 + 4 * lane;
        uint8_t* dst = p.out + i * 512 + 4 * lane;                      // the output is contiguous: plane * HW + chunk * 512
        for (; i < i_end; ++i) {
            float v[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) Vec4<T>::load(src + 128 * k, v[k]);
            uint32_t by[4][4];
            uint32_t redo = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bool d;
                    by[k][j] = fast_byte<MODE>(v[k][j], p, d);
                    if (MODE >= kModeSig) redo |= d ? (1u << (4 * k + j)) : 0u;
                }
            if (MODE >= kModeSig && redo) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (redo & (1u << (4 * k + j))) by[k][j] = exact_byte<MODE>(v[k][j], p.thr);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t wd = by[k][0] | (by[k][1] << 8) | (by[k][2] << 16) | (by[k][3] << 24);
                asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(dst + 128 * k), "r"(wd) : "memory");
            }
            dst += 512;
            src += 512;
            if (++chunk == cpp) {   // next plane: step over the strides
                chunk = 0;
                src -= p.HW;
                if (++c == p.C) { c = 0; src += p.sn - (int64_t)(p.C - 1) * p.sc; }
                else src += p.sc;
            }
        }
    } else {
        const int64_t total = (int64_t)p.N * p.C * p.HW;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t plane = i / p.HW, e = i - plane * p.HW;
            const int64_t n = plane / p.C, c = plane - n * p.C;
            const float v = Vec4<T>::load1(reinterpret_cast<const T*>(p.x) + n * p.sn + c * p.sc + e);
            bool d;
            uint32_t b = fast_byte<MODE>(v, p, d);
            if (MODE >= kModeSig && d) b = exact_byte<MODE>(v, p.thr);
            p.out[i] = (uint8_t)b;
        }
    }
}

template <typename T, int VEC>
static void launch_masks(const MaskArgs& p, int mode, unsigned grid, cudaStream_t st) {
    switch (mode) {
        case kModeProb: masks_u8_kernel<T, VEC, kModeProb><<<grid, 256, 0, st>>>(p); break;
        case kModeProbThr: masks_u8_kernel<T, VEC, kModeProbThr><<<grid, 256, 0, st>>>(p); break;
        case kModeSig: masks_u8_kernel<T, VEC, kModeSig><<<grid, 256, 0, st>>>(p); break;
        default: masks_u8_kernel<T, VEC, kModeSigThr><<<grid, 256, 0, st>>>(p); break;
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
    const bool v16 = HW % 512 == 0 && x->sn % 4 == 0 && x->sc % 4 == 0 && reinterpret_cast<uintptr_t>(x->ptr) % (4 * esz) == 0 &&
                     reinterpret_cast<uintptr_t>(out) % 4 == 0;
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -10;
    const int64_t total = v16 ? (int64_t)N * C * (HW / 512) * 32 : (int64_t)N * C * HW;   // threads wanted
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;   // resident wave, grid-stride beyond it
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int mode = (x_is_prob ? 0 : 2) + (use_threshold ? 1 : 0);
    if (x->dtype == ECO_F32) {
        if (v16) launch_masks<float, 16>(p, mode, (unsigned)grid, st);
        else launch_masks<float, 1>(p, mode, (unsigned)grid, st);
    } else {
        if (v16) launch_masks<__nv_bfloat16, 16>(p, mode, (unsigned)grid, st);
        else launch_masks<__nv_bfloat16, 1>(p, mode, (unsigned)grid, st);
    }
    return check_cuda(cudaGetLastError(), "masks_u8_kernel launch");
}
