// Shared device helpers for the loss / scoring kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ecoloss.h"

namespace eco {

constexpr float kEps = 1e-7f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr double kLn2d = 0.6931471805599453094;

// stat indices (mirrors oracle/closed_form.py)
enum : int { S_N = 0, S_A = 1, S_B = 2, S_AB = 3, S_BB = 4, S_SP = 5, S_FL = 6, S_FLB = 7 };

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define ECO_CUDA(expr)                                        \
    do {                                                      \
        int _rc = ::eco::check_cuda((expr), #expr);           \
        if (_rc) return _rc;                                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int sm_count_cached(int device);

// ---------------------------------------------------------------------------------------------
// streaming loads / stores (read-once data: bypass L1 allocation)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// 4 consecutive elements of dtype T starting at p (16-byte aligned for float, 8-byte for bf16).
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 r = ldg_stream_f4(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        stg_stream_f4(p, make_float4(v[0], v[1], v[2], v[3]));
    }
    static __device__ __forceinline__ float load1(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        uint2 r = ldg_stream_u2(p);
        v[0] = __uint_as_float(r.x << 16);
        v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16);
        v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&lo);
        r.y = *reinterpret_cast<uint32_t*>(&hi);
        stg_stream_u2(p, r);
    }
    static __device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// uint8 masks (labels stored as bytes: 1 B instead of 4 B per element)
template <>
struct Vec4<uint8_t> {
    static __device__ __forceinline__ void load(const uint8_t* p, float (&v)[4]) {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
        // byte k -> float without I2F: 0x4b000000 | byte is 8388608 + byte exactly
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __uint_as_float(0x4b000000u | ((r >> (8 * k)) & 0xffu)) - 8388608.0f;
    }
    static __device__ __forceinline__ float load1(const uint8_t* p) { return (float)__ldg(p); }
};

// ---------------------------------------------------------------------------------------------
// per-element math
// ---------------------------------------------------------------------------------------------
// Bit-compatible with ATen's CUDA sigmoid for fp32: 1 / (1 + expf(-x)), IEEE division, no fast-math.
// (thresholded counts and the |x_i - x_j| kink need the same bits as torch.sigmoid on this device.)
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// log2(1 + exp(-|b|)); multiply the SUM by ln2 afterwards.  These terms only ever enter sums of
// >= thousands of elements compared at 1e-5, so the 2-ulp MUFU approximations are ample.
__device__ __forceinline__ float softplus_neg_abs_log2(float b) { return lg2_approx(1.0f + ex2_approx(-fabsf(b) * kLog2e)); }

// (1-b)^1.5 * log2(b + eps)  (NaN for b > 1 or b < -eps, like the reference's torch.pow/log).
__device__ __forceinline__ float focal_fg_log2(float b) {
    float t = 1.0f - b;
    return (t * sqrt_approx(t)) * lg2_approx(b + kEps);
}
// b^1.5 * log2(1 - b + eps)
__device__ __forceinline__ float focal_bg_log2(float b) { return (b * sqrt_approx(b)) * lg2_approx((1.0f - b) + kEps); }

__device__ __forceinline__ float sigmoid_fast(float b) { return rcp_approx(1.0f + ex2_approx(-b * kLog2e)); }

// The BCE terms see probabilities and labels, i.e. arguments in [0, 1], where softplus and sigmoid are polynomials in
// t = b^2 (both remainders are even / odd functions, so the fits hold on [-1, 1]):
//   softplus(-b) = ln2 - b/2 + t/8 + t^2 r(t)   (max abs err 7e-9)      sigmoid(b) = 1/2 + b s(t)   (max abs err 7e-8)
constexpr float kSpR0 = -5.2077806842e-03f, kSpR1 = 3.4455654967e-04f, kSpR2 = -2.2275527791e-05f;
constexpr float kSgS0 = 2.4999950727e-01f, kSgS1 = -2.0825986369e-02f, kSgS2 = 2.054964847e-03f,
                kSgS3 = -1.6997672007e-04f;
// softplus(b) = max(b, 0) + log(1 + exp(-|b|)) of an arbitrary slot value: the polynomial inside [-1, 1] (two MUFU operations
// per element less: the pair-leaf statistics kernel is XU-limited), the MUFU form outside.  (The same switch for the sigmoid
// of the gradient kernel bought nothing: 93.8 against 92.8 us at 54x3x512.)
__device__ __forceinline__ float softplus_slot(float b) {
    if (fabsf(b) <= 1.0f) {
        const float t = b * b;
        float r = fmaf(kSpR2, t, kSpR1);
        r = fmaf(r, t, kSpR0);
        return fmaf(t * t, r, fmaf(0.125f, t, fmaf(0.5f, b, kLn2)));
    }
    return fmaf(softplus_neg_abs_log2(b), kLn2, fmaxf(b, 0.f));
}

// d/db [ -(1-b)^1.5 log(b+eps) ] = 1.5 sqrt(1-b) log(b+eps) - (1-b)^1.5/(b+eps)
__device__ __forceinline__ float dfocal_fg(float b) {
    float t = 1.0f - b;
    float s = sqrt_approx(t);
    float be = b + kEps;
    return 1.5f * s * (lg2_approx(be) * kLn2) - (t * s) * rcp_approx(be);
}
// d/db [ -b^1.5 log(1-b+eps) ] = -1.5 sqrt(b) log(1-b+eps) + b^1.5/(1-b+eps)
__device__ __forceinline__ float dfocal_bg(float b) {
    float s = sqrt_approx(b);
    float te = (1.0f - b) + kEps;
    return -1.5f * s * (lg2_approx(te) * kLn2) + (b * s) * rcp_approx(te);
}

// Non-default focal exponent (focal_loss(gamma=...), loss_functions.py:46-48): the same terms through powf, which
// is what torch.pow lowers to on this device (so negative bases / integer exponents behave as in the reference).
// Off the hot path: no caller in the reference passes gamma, the kernels take this branch only when asked.
__device__ __forceinline__ float focal_fg_log2_gen(float b, float gamma) { return powf(1.0f - b, gamma) * lg2_approx(b + kEps); }
__device__ __forceinline__ float focal_bg_log2_gen(float b, float gamma) { return powf(b, gamma) * lg2_approx((1.0f - b) + kEps); }
__device__ __forceinline__ float dfocal_fg_gen(float b, float gamma) {
    const float t = 1.0f - b, be = b + kEps;
    return gamma * powf(t, gamma - 1.0f) * (lg2_approx(be) * kLn2) - powf(t, gamma) * rcp_approx(be);
}
__device__ __forceinline__ float dfocal_bg_gen(float b, float gamma) {
    const float te = (1.0f - b) + kEps;
    return -gamma * powf(b, gamma - 1.0f) * (lg2_approx(te) * kLn2) + powf(b, gamma) * rcp_approx(te);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// closed forms: one leaf's sums -> 7 losses and their Jacobian w.r.t. stats [1..7]
// (float64; mirrors oracle/closed_form.py:leaf_losses / leaf_coefs; reference lines cited there)
// ---------------------------------------------------------------------------------------------
struct LeafOut {
    double loss[ECO_NLOSS];
    double jac[ECO_NLOSS][ECO_NJAC];  // d loss_k / d (Sa, Sb, Sab, Sbb, SP, FL, FLB)
};

// phi(t) = -(1-t)^1.8 log(t+eps) and its derivative (focal_dice_coefficient, loss_functions.py:101).
// `omt` = 1-t and `te` = t+eps are formed in float64 (1-t cancels when the Dice coefficient is near 1); the
// transcendentals themselves run in fp32 (rel. error ~1e-7, two orders inside the 1e-5 budget) because a
// float64 pow/log chain here sits on the critical path between the two passes of the fused kernel.
__device__ __forceinline__ void phi_fd(double omt, double te, double& phi, double& dphi) {
    const float l = logf((float)te);
    const float p08 = powf((float)omt, 0.8f);
    const double p18 = (double)p08 * omt;
    phi = -p18 * (double)l;
    dphi = 1.8 * (double)p08 * (double)l - p18 / te;
}
// same with a caller-chosen exponent (focal_dice_coefficient(gamma=...)); float64 throughout, off the hot path
__device__ inline void phi_fd_gen(double omt, double te, double gamma, double& phi, double& dphi) {
    const double l = log(te);
    const double pg = pow(omt, gamma);
    phi = -pg * l;
    dphi = gamma * pow(omt, gamma - 1.0) * l - pg / te;
}

// Shape parameters of the Tversky and focal-Dice closed forms (loss_functions.py:82,96); the fused kernels use the
// reference's defaults at compile time, the stand-alone primitives may pass their own.
struct LeafShape {
    double alpha = 0.5, beta = 0.3, fd_gamma = 1.8;
};

struct LeafMoments {
    double n, I, D, Ib, Db, FN, FP;
};
__device__ __forceinline__ LeafMoments leaf_moments(const double* s) {
    LeafMoments m;
    m.n = s[S_N];
    m.I = s[S_AB];
    m.D = s[S_A] + s[S_BB];
    m.Ib = m.n - s[S_A] - s[S_B] + s[S_AB];
    m.Db = (m.n - s[S_A]) + (m.n - 2 * s[S_B] + s[S_BB]);
    m.FN = s[S_A] - s[S_AB];
    m.FP = s[S_B] - s[S_AB];
    return m;
}

// One of the 7 losses of a leaf (k = 0..6) and its row of the Jacobian, both times `scale`.
// Rows are independent, so the fused kernel spreads them over threads.
template <bool kDefaultShape = true>
__device__ inline void leaf_closed_form_row(const double* s, double bw, double scale, int k, double& loss,
                                            double (&jrow)[ECO_NJAC], const LeafShape& shape = LeafShape()) {
    const double eps = 1e-7, m = 10 * 0.33;
    const double alpha = kDefaultShape ? 0.5 : shape.alpha, beta = kDefaultShape ? 0.3 : shape.beta;
    const LeafMoments M = leaf_moments(s);
    const double n = M.n, I = M.I, D = M.D, Ib = M.Ib, Db = M.Db, FN = M.FN, FP = M.FP;
#pragma unroll
    for (int j = 0; j < ECO_NJAC; ++j) jrow[j] = 0.0;
    double dI = 0, dD = 0, dIb = 0, dDb = 0, dFN = 0, dFP = 0;
    loss = 0.0;
    switch (k) {
        case 1:  // bce
            loss = (s[S_SP] - s[S_AB]) / n;
            jrow[2] = -1.0 / n;
            jrow[4] = 1.0 / n;
            break;
        case 2:  // focal
            loss = (s[S_FL] + bw * s[S_FLB]) / n;
            jrow[5] = 1.0 / n;
            jrow[6] = bw / n;
            break;
        case 3: {  // dice
            const double r = 1.0 / (D + eps), rb = 1.0 / (2 * Db + eps);
            loss = m * (-(2 * I + eps) * r - bw * (2 * Ib + eps) * rb);
            dI = -m * 2 * r;
            dD = m * (2 * I + eps) * r * r;
            dIb = -m * bw * 2 * rb;
            dDb = m * bw * (2 * Ib + eps) * 2 * rb * rb;
            break;
        }
        case 4: {  // generalized dice
            const double r = 1.0 / (D + eps), rb = 1.0 / (Db + eps);
            loss = m * (-((I + eps) * r + bw * (Ib + eps) * rb));
            dI = -m * r;
            dD = m * (I + eps) * r * r;
            dIb = -m * bw * rb;
            dDb = m * bw * (Ib + eps) * rb * rb;
            break;
        }
        case 5: {  // tversky
            const double r1 = 1.0 / (I + alpha * FN + beta * FP + eps);
            const double r2 = 1.0 / (Ib + alpha * FP + beta * FN + eps);
            loss = m * (-(I + eps) * r1 + bw * (-(Ib + eps) * r2));
            // -1/t1 + (I+eps)/t1^2 = -(alpha FN + beta FP)/t1^2: no cancellation
            dI = -m * (alpha * FN + beta * FP) * r1 * r1;
            dIb = -m * bw * (alpha * FP + beta * FN) * r2 * r2;
            dFN = m * ((I + eps) * alpha * r1 * r1 + bw * (Ib + eps) * beta * r2 * r2);
            dFP = m * ((I + eps) * beta * r1 * r1 + bw * (Ib + eps) * alpha * r2 * r2);
            break;
        }
        case 6: {  // focal dice
            const double r = 1.0 / (D + eps);
            const double dc = (2 * I + eps) * r;
            double phi, dphi;
            if constexpr (kDefaultShape) phi_fd((D - 2 * I) * r, dc + eps, phi, dphi);  // 1 - dc = (D - 2I)/(D + eps)
            else phi_fd_gen((D - 2 * I) * r, dc + eps, shape.fd_gamma, phi, dphi);
            loss = m * phi;
            dI = m * dphi * 2 * r;
            dD = -m * dphi * dc * r;
            if (bw != 0.0) {
                const double rb = 1.0 / (Db + eps);
                const double dcb = (2 * Ib + eps) * rb;
                double phib, dphib;
                if constexpr (kDefaultShape) phi_fd((Db - 2 * Ib) * rb, dcb + eps, phib, dphib);
                else phi_fd_gen((Db - 2 * Ib) * rb, dcb + eps, shape.fd_gamma, phib, dphib);
                loss += m * bw * phib;
                dIb = m * bw * dphib * 2 * rb;
                dDb = -m * bw * dphib * dcb * rb;
            }
            break;
        }
        default:
            break;
    }
    if (k >= 3) {
        // chain to the sums: Sa: +D -Ib -Db +FN ; Sb: -Ib -2Db +FP ; Sab: +I +Ib -FN -FP ; Sbb: +D +Db
        jrow[0] = dD - dIb - dDb + dFN;
        jrow[1] = -dIb - 2 * dDb + dFP;
        jrow[2] = dI + dIb - dFN - dFP;
        jrow[3] = dD + dDb;
    }
    loss *= scale;
#pragma unroll
    for (int j = 0; j < ECO_NJAC; ++j) jrow[j] *= scale;
}

__device__ inline void leaf_closed_form(const double* s, double bw, double scale, LeafOut& o) {
    for (int k = 0; k < ECO_NLOSS; ++k) leaf_closed_form_row(s, bw, scale, k, o.loss[k], o.jac[k]);
}
__device__ inline void leaf_closed_form(const double* s, double bw, double scale, const LeafShape& shape, LeafOut& o) {
    for (int k = 0; k < ECO_NLOSS; ++k) leaf_closed_form_row<false>(s, bw, scale, k, o.loss[k], o.jac[k], shape);
}

// coefficient vector c[j] = sum_k upstream[k] * jac[k][j]  (7 values), as floats for the per-pixel pass.
struct LeafCoef {
    float sa, sb, sab, sbb2 /* = 2*c_Sbb */, sp, fl, flb;
};

__device__ __forceinline__ LeafCoef make_coef(const double* jac /*[7][7]*/, const float* upstream) {
    double c[ECO_NJAC];
#pragma unroll
    for (int j = 0; j < ECO_NJAC; ++j) c[j] = 0.0;
    for (int k = 1; k < ECO_NLOSS; ++k) {
        const double w = (double)upstream[k];
        if (w != 0.0) {
#pragma unroll
            for (int j = 0; j < ECO_NJAC; ++j) c[j] += w * jac[k * ECO_NJAC + j];
        }
    }
    LeafCoef r;
    r.sa = (float)c[0];
    r.sb = (float)c[1];
    r.sab = (float)c[2];
    r.sbb2 = (float)(2.0 * c[3]);
    r.sp = (float)c[4];
    r.fl = (float)c[5];
    r.flb = (float)c[6];
    return r;
}

}  // namespace eco
