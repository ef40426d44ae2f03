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

__device__ __forceinline__ double phi_fd(double t) { return -pow(1.0 - t, 1.8) * log(t + 1e-7); }
__device__ __forceinline__ double dphi_fd(double t) {
    return 1.8 * pow(1.0 - t, 0.8) * log(t + 1e-7) - pow(1.0 - t, 1.8) / (t + 1e-7);
}

__device__ inline void leaf_closed_form(const double* s, double bw, double scale, LeafOut& o) {
    const double eps = 1e-7, m = 10 * 0.33, alpha = 0.5, beta = 0.3;
    const double n = s[S_N];
    const double I = s[S_AB], D = s[S_A] + s[S_BB];
    const double Ib = n - s[S_A] - s[S_B] + s[S_AB];
    const double Db = (n - s[S_A]) + (n - 2 * s[S_B] + s[S_BB]);
    const double FN = s[S_A] - s[S_AB], FP = s[S_B] - s[S_AB];
    const double t1 = I + alpha * FN + beta * FP + eps;
    const double t2 = Ib + alpha * FP + beta * FN + eps;
    const double dc = (2 * I + eps) / (D + eps), dcb = (2 * Ib + eps) / (Db + eps);

    for (int k = 0; k < ECO_NLOSS; ++k)
        for (int j = 0; j < ECO_NJAC; ++j) o.jac[k][j] = 0.0;

    o.loss[0] = 0.0;
    o.loss[1] = (s[S_SP] - s[S_AB]) / n;
    o.loss[2] = (s[S_FL] + bw * s[S_FLB]) / n;
    o.loss[3] = m * (-(2 * I + eps) / (D + eps) - bw * (2 * Ib + eps) / (2 * Db + eps));
    o.loss[4] = m * (-((I + eps) / (D + eps) + bw * (Ib + eps) / (Db + eps)));
    o.loss[5] = m * (-(I + eps) / t1 + bw * (-(Ib + eps) / t2));
    const double phi_b = (bw != 0.0) ? phi_fd(dcb) : 0.0;
    o.loss[6] = m * (phi_fd(dc) + bw * phi_b);

    // partials w.r.t. the intermediate moments (I, D, Ib, Db, FN, FP), then chained to the sums.
    // chain: Sa: +D -Ib -Db +FN ; Sb: -Ib -2Db +FP ; Sab: +I +Ib -FN -FP ; Sbb: +D +Db
    auto chain = [&](int k, double dI, double dD, double dIb, double dDb, double dFN, double dFP) {
        o.jac[k][0] = dD - dIb - dDb + dFN;
        o.jac[k][1] = -dIb - 2 * dDb + dFP;
        o.jac[k][2] = dI + dIb - dFN - dFP;
        o.jac[k][3] = dD + dDb;
    };
    // bce
    o.jac[1][2] = -1.0 / n;
    o.jac[1][4] = 1.0 / n;
    // focal
    o.jac[2][5] = 1.0 / n;
    o.jac[2][6] = bw / n;
    // dice
    chain(3, -m * 2 / (D + eps), m * (2 * I + eps) / ((D + eps) * (D + eps)), -m * bw * 2 / (2 * Db + eps),
          m * bw * (2 * Ib + eps) * 2 / ((2 * Db + eps) * (2 * Db + eps)), 0.0, 0.0);
    // generalized dice
    chain(4, -m / (D + eps), m * (I + eps) / ((D + eps) * (D + eps)), -m * bw / (Db + eps),
          m * bw * (Ib + eps) / ((Db + eps) * (Db + eps)), 0.0, 0.0);
    // tversky
    chain(5, m * (-1 / t1 + (I + eps) / (t1 * t1)), 0.0, m * bw * (-1 / t2 + (Ib + eps) / (t2 * t2)), 0.0,
          m * ((I + eps) * alpha / (t1 * t1) + bw * (Ib + eps) * beta / (t2 * t2)),
          m * ((I + eps) * beta / (t1 * t1) + bw * (Ib + eps) * alpha / (t2 * t2)));
    // focal dice
    {
        const double p1 = dphi_fd(dc);
        const double p2 = (bw != 0.0) ? dphi_fd(dcb) : 0.0;
        chain(6, m * p1 * 2 / (D + eps), -m * p1 * (2 * I + eps) / ((D + eps) * (D + eps)),
              m * bw * p2 * 2 / (Db + eps), -m * bw * p2 * (2 * Ib + eps) / ((Db + eps) * (Db + eps)), 0.0, 0.0);
    }
    for (int k = 0; k < ECO_NLOSS; ++k) {
        o.loss[k] *= scale;
        for (int j = 0; j < ECO_NJAC; ++j) o.jac[k][j] *= scale;
    }
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
