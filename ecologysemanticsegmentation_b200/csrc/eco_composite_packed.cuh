// Packed-fp32x2 (FFMA2/FADD2/FMUL2) implementation of the two SEPARATE passes of the 3-organ composite loss for the aligned
// 128-bit path: composite3_stats_packed_kernel / composite3_grad_packed_kernel (statistics -> all-reduce by NCCL -> gradient,
// and the autograd path on inputs the one-launch kernel does not serve).  Included by eco_composite.cu after the shared
// layout / helper definitions.  (The one-launch kernel of this generation was removed in round 2: eco_composite_v3.cuh
// serves fp32 / bf16 logits and fp32 probabilities, the scalar kernel everything else.)
//
// Why this shape (measured on B200, profiles/microbench/pipes.cu): the FMA pipe sustains 128 fp32 lanes
// /clk/SM with scalar FFMA *or* with FFMA2, but FFMA2 needs half the issue slots, and MUFU (16 lanes/clk/SM)
// overlaps fully with an FFMA2 stream.  The composite loss is ~420 FMA-class ops and ~40 MUFU per pixel, far
// above what 12 B/element of HBM traffic can hide, so the kernel is built to keep the FMA and XU pipes busy:
//   * two pixels per instruction (float2 lanes);
//   * sigmoid = ex2.approx + rcp.approx; the exact (ATen-bit-compatible) sigmoid is recomputed only where
//     |p_i - p_j| < 4e-6, i.e. where the sign of the |.| kink could depend on the last bits;
//   * on the from-logits path every b operand lies in [0,1], so log(1+exp(-b)) and sigmoid(b) are short
//     even polynomials in b^2 on the FMA pipe instead of two MUFU each; the linear and quadratic parts of the
//     softplus series fold into sums that are accumulated anyway (sum b, sum b^2);
//   * u(sp, p) = sp + p * (0.5 - 0.5 sp): one FFMA2, and sum u follows algebraically from other sums;
//   * pass 1 is role-split: warps 0-3 / 4-7 / 8-11 of a 384-thread CTA each own one organ pair (26 packed
//     accumulators instead of 84), all three roles walk the same pixels so the second read of a plane hits L1;
//   * pass 2 walks each CTA's range backwards so that it starts on the lines pass 1 left in L2.
#pragma once

namespace eco {

constexpr int kPThreads = 384;
constexpr int kPWarps = kPThreads / 32;
constexpr int kRoleWarps = kPWarps / 3;
constexpr int kRoleThreads = kRoleWarps * 32;
constexpr int kRAcc = 29;
constexpr int kPFlushIters = 8;
constexpr float kTieEps = 4e-6f;

// per-role accumulator indices
enum : int {
    R_G = 0, R_GD, R_DS, R_X, R_XX, R_GX, R_RX, R_FLX, R_M1, R_M1G, R_M2, R_M2G, R_M3, R_M3G,
    R_U = 14,  // + 4k + {0 UU, 1 GU, 2 RU, 3 FLU}, k = 0..2
    R_GJ = 26, R_GGJ = 27, R_GDD = 28  // sum g_j, sum g_j^2, sum gd^2: label-b second moments (binary <=> GG == G)
};

typedef float2 f2;
__device__ __forceinline__ f2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 abs2(f2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

// (kSpR0..2, kSgS0..3: the softplus / sigmoid polynomials on [-1, 1], eco_common.cuh)

__device__ __forceinline__ f2 sigmoid_fast2(f2 z) {
    const f2 t = mul2(z, splat(-kLog2e));
    const f2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
    const f2 d = add2(e, splat(1.0f));
    return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}

// accumulate the b-only terms of a leaf: the softplus remainder (or the whole term) and the focal term
template <bool UNIT>
__device__ __forceinline__ void leaf_b_terms2(f2 b, f2 t, f2& r_acc, f2& fl_acc) {
    if (UNIT) {
        const f2 t2 = mul2(t, t);
        f2 r = fma2(splat(kSpR2), t, splat(kSpR1));
        r = fma2(r, t, splat(kSpR0));
        r_acc = fma2(t2, r, r_acc);
    } else {
        r_acc.x += fmaf(softplus_neg_abs_log2(b.x), kLn2, fmaxf(b.x, 0.f));
        r_acc.y += fmaf(softplus_neg_abs_log2(b.y), kLn2, fmaxf(b.y, 0.f));
    }
    const f2 om = fma2(b, splat(-1.0f), splat(1.0f));
    const f2 s = make_float2(sqrt_approx(om.x), sqrt_approx(om.y));
    const f2 w = mul2(om, s);
    const f2 be = add2(b, splat(kEps));
    const f2 l = make_float2(lg2_approx(be.x), lg2_approx(be.y));
    fl_acc = fma2(w, l, fl_acc);
}

// one pixel pair of one organ pair (i, j); `pc`/`gc` = the channel whose plain leaf this role owns
template <bool UNIT>
__device__ __forceinline__ void role_plain_stats(f2 pc, f2 gc, f2 (&acc)[kRAcc]) {
    // plain leaf (a = g_c, b = p_c)
    const f2 t = mul2(pc, pc);
    acc[R_G] = add2(acc[R_G], gc);
    acc[R_X] = add2(acc[R_X], pc);
    acc[R_XX] = add2(acc[R_XX], t);
    acc[R_GX] = fma2(gc, pc, acc[R_GX]);
    leaf_b_terms2<UNIT>(pc, t, acc[R_RX], acc[R_FLX]);
}

template <bool UNIT>
__device__ __forceinline__ void role_pair_stats(f2 pi, f2 pj, f2 gi, f2 gj, f2 d, f2 (&acc)[kRAcc]) {
    const f2 hh = fma2(pi, splat(-0.5f), splat(0.5f));
    const f2 m1 = mul2(pi, pj);
    const f2 q = mul2(pi, d);
    const f2 m3 = mul2(pi, q);
    const f2 u1 = fma2(pj, hh, pi);
    const f2 u2 = fma2(d, hh, pi);
    const f2 u3 = fma2(q, hh, pi);
    const f2 gd = abs2(fma2(gj, splat(-1.0f), gi));
    acc[R_GD] = add2(acc[R_GD], gd);
    acc[R_DS] = add2(acc[R_DS], d);
    acc[R_GJ] = add2(acc[R_GJ], gj);
    acc[R_GGJ] = fma2(gj, gj, acc[R_GGJ]);
    acc[R_GDD] = fma2(gd, gd, acc[R_GDD]);
    // intersection leaves (a = m_k, b = label)
    acc[R_M1] = add2(acc[R_M1], m1);
    acc[R_M1G] = fma2(m1, gj, acc[R_M1G]);
    acc[R_M2] = add2(acc[R_M2], q);
    acc[R_M2G] = fma2(q, gd, acc[R_M2G]);
    acc[R_M3] = add2(acc[R_M3], m3);
    acc[R_M3G] = fma2(m3, gd, acc[R_M3G]);
    // union leaves (a = g_i, b = u_k)
    const f2 us[3] = {u1, u2, u3};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const f2 u = us[k];
        const f2 t = mul2(u, u);
        acc[R_U + 4 * k + 0] = add2(acc[R_U + 4 * k + 0], t);
        acc[R_U + 4 * k + 1] = fma2(gi, u, acc[R_U + 4 * k + 1]);
        leaf_b_terms2<UNIT>(u, t, acc[R_U + 4 * k + 2], acc[R_U + 4 * k + 3]);
    }
}

__device__ __forceinline__ bool flush_role_acc(f2 (&acc)[kRAcc], double* warp_slot /* smem [32] */, int lane) {
    const bool nonbinary = acc[R_GGJ].x != acc[R_GJ].x || acc[R_GGJ].y != acc[R_GJ].y ||
                           acc[R_GDD].x != acc[R_GD].x || acc[R_GDD].y != acc[R_GD].y;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < kRAcc) ? acc[i].x + acc[i].y : 0.f;
    const float tot = butterfly32(v, lane);
    warp_slot[lane] += (double)tot;
#pragma unroll
    for (int k = 0; k < kRAcc; ++k) acc[k] = splat(0.f);
    return nonbinary;
}

// Arguments of the in-kernel all-reduce over NVLink peer memory (the sharded fused step, eco_composite_v3.cuh).  Every rank
// owns an exchange buffer mapped into all peers (CUDA IPC).  The buffer starts with the slots of the first-generation
// protocol (double slots[2 parities][world][128], then u32 flags[world]; the kernel that used them was removed in round 2,
// the layout is kept so that buffers stay interchangeable); the LL rows of the current protocol follow.
struct XchArgs {
    double* const* peers;  // device array [world] of peer-mapped exchange buffers ([rank] = own)
    int rank, world;
    unsigned int epoch;
    unsigned int* status;  // device-visible word (workspace or mapped pinned host memory): set to 1 if a wait timed out
    unsigned long long timeout_ns;  // v3 kernels: wall-clock limit of one wait on a peer
};

__host__ __device__ inline size_t xch_flags_offset_doubles(int world) { return (size_t)2 * world * 128; }

struct PStatsSmem {
    double warp_slots[kPWarps][32];
    double role_sums[3][32];
    double corr[15];
    bool is_last;
};

// organ channel -> role that owns its plain leaf; role -> (i, j, c)
__host__ __device__ constexpr int role_of_channel(int c) { return c == 0 ? 1 : (c == 1 ? 0 : 2); }
__host__ __device__ constexpr int role_channel(int r) { return r == 0 ? 1 : (r == 1 ? 0 : 2); }

// block-level conversion of the three roles' sums into the shared 100-slot layout (see eco_composite.cu)
template <bool UNIT>
__device__ inline double packed_to_layout(const double (*R)[32], const double* corr, int idx, double n_blk) {
    auto X = [&](int c) { return R[role_of_channel(c)][R_X]; };
    auto sp_of = [&](double sb, double sbb, double r) { return UNIT ? n_blk * kLn2d + 0.5 * sb + 0.125 * sbb + r : r; };
    if (idx == A_N) return n_blk;
    if (idx < A_GD) return R[role_of_channel(idx - A_G)][R_G];
    if (idx < A_CH) return R[idx - A_GD][R_GD];
    if (idx < A_PAIR) {
        const int c = (idx - A_CH) / 5, k = (idx - A_CH) % 5;
        const double* r = R[role_of_channel(c)];
        switch (k) {
            case 0: return r[R_X];
            case 1: return r[R_XX];
            case 2: return r[R_GX];
            case 3: return sp_of(r[R_X], r[R_XX], r[R_RX]);
            default: return -kLn2d * r[R_FLX];
        }
    }
    if (idx < A_CORR) {
        const int p = (idx - A_PAIR) / 21, k = (idx - A_PAIR) % 21;
        const double* r = R[p];
        const int i = pair_i(p), j = pair_j(p);
        const int grp = k / 7, kk = k % 7;  // grp 0: (M1, U1) 1: (M2, U2) 2: (M3, U3)
        const double m = r[R_M1 + 2 * grp], mg = r[R_M1G + 2 * grp];
        // sum u_k = X_i + 0.5 (sum p_k - sum x_i p_k): p_1 = x_j, p_2 = d, p_3 = q = x_i d
        const double psum = grp == 0 ? X(j) : (grp == 1 ? r[R_DS] : r[R_M2]);
        const double usum = X(i) + 0.5 * (psum - m);
        const double* ul = r + R_U + 4 * grp;
        switch (kk) {
            case 0: return m;
            case 1: return mg;
            case 2: return usum;
            case 3: return ul[0];
            case 4: return ul[1];
            case 5: return sp_of(usum, ul[0], ul[2]);
            default: return -kLn2d * ul[3];
        }
    }
    // label-b corrections: sum(b^2 - b) follows from the second-moment accumulators, the transcendental
    // corrections from the (rare) slow pass over non-binary labels
    const int L = (idx - A_CORR) / 3, k = (idx - A_CORR) % 3;
    if (k != 0) return corr[idx - A_CORR];
    if (L < 2) return R[L][R_GGJ] - R[L][R_GJ];          // g1 via role 0 (j = 1), g2 via role 1 (j = 2)
    return R[L - 2][R_GDD] - R[L - 2][R_GD];              // gd of pair L-2
}

// default-cached 128-bit load (the three roles of a CTA read each plane twice: keep it in L1)
template <typename T>
__device__ __forceinline__ void ld4_cached(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4_cached<float>(const float* p, float (&v)[4]) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <>
__device__ __forceinline__ void ld4_cached<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// cp.async (LDGSTS): BYTES = 16 (4 x f32) or 8 (4 x bf16), L1-allocating so the second role's copy of a plane hits
// streaming flavour (L2 only; 16-byte copies only): pass 2 reads every line once
template <int BYTES>
__device__ __forceinline__ void cp_async_cg(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__device__ __forceinline__ void lds_x4(const void* p, float (&v)[4]);
template <>
__device__ __forceinline__ void lds_x4<float>(const void* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <>
__device__ __forceinline__ void lds_x4<__nv_bfloat16>(const void* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xffff0000u);
}

// 32-bit shared-window addressing: generic char* arithmetic costs several 64-bit IMADs per access
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES, bool STREAM>
__device__ __forceinline__ void cp_async_s(uint32_t dst, const void* gmem_src) {
    if (STREAM && BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// staged unit of x (4 elements) at shared address a -> 4 floats
template <typename TX>
__device__ __forceinline__ float4 lds_unit(uint32_t a);
template <>
__device__ __forceinline__ float4 lds_unit<float>(uint32_t a) { return lds_f4(a); }
template <>
__device__ __forceinline__ float4 lds_unit<__nv_bfloat16>(uint32_t a) {
    const uint2 r = lds_u2(a);
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
}

// running global cursor over the float4 units of one CTA range for a fixed thread stride: byte offsets of the
// current unit inside plane 0 of x and of g, advanced without multiplications (one compare per step)
struct UnitCursor {
    int64_t q, off, upp, n;
    int64_t xoff, goff;          // element offsets n * sn + off * 4
    int64_t x_sn, g_sn;
    __device__ __forceinline__ void init(int64_t q0, const CompArgs& a) {
        q = q0; upp = a.units_per_plane; x_sn = a.x_sn; g_sn = a.g_sn;
        n = q / upp;
        off = q - n * upp;
        xoff = n * x_sn + off * 4;
        goff = n * g_sn + off * 4;
    }
    __device__ __forceinline__ void advance(int stride) {
        q += stride; off += stride; xoff += (int64_t)stride * 4; goff += (int64_t)stride * 4;
        while (off >= upp) {   // next image: planes of one image are sn apart, not upp * 4
            off -= upp; ++n;
            xoff += x_sn - upp * 4;
            goff += g_sn - upp * 4;
        }
    }
    __device__ __forceinline__ void retreat(int stride) {
        q -= stride; off -= stride; xoff -= (int64_t)stride * 4; goff -= (int64_t)stride * 4;
        while (off < 0) {
            off += upp; --n;
            xoff -= x_sn - upp * 4;
            goff -= g_sn - upp * 4;
        }
    }
};

// dynamic shared memory of the packed kernels: 2 stages x 6 planes x kPThreads x 16 B (pass 2 uses all 6 planes)
constexpr int kStageBytes = 2 * 6 * kPThreads * 16;

// walks the float4 units [lo, hi) of a CTA with a fixed thread stride, tracking (image, offset) incrementally
struct UnitWalker {
    int64_t q, n, off, upp;
    __device__ __forceinline__ void init(int64_t q0, int64_t units_per_plane) {
        q = q0; upp = units_per_plane;
        n = q / upp;
        off = q - n * upp;
    }
    __device__ __forceinline__ void advance(int stride) {
        q += stride;
        off += stride;
        while (off >= upp) { off -= upp; ++n; }
    }
};

// Pass 1 (packed, role-split).  Same contract as stats_phase(): the LAST CTA leaves acc_out[0..100).
template <typename TX, bool LOGITS>
__device__ __forceinline__ void stats_phase_packed(const CompArgs& a, PStatsSmem& sm, char* stage_smem,
                                                   unsigned int* __restrict__ counter, double* __restrict__ partials,
                                                   double* __restrict__ acc_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = warp / kRoleWarps;
    const int rtid = threadIdx.x - role * kRoleThreads;
    for (int i = threadIdx.x; i < kPWarps * 32; i += kPThreads) (&sm.warp_slots[0][0])[i] = 0.0;
    if (threadIdx.x < 15) sm.corr[threadIdx.x] = 0.0;
    __syncthreads();

    const int ci = pair_i(role), cj = pair_j(role);
    const bool chan_is_i = role_channel(role) == ci;
    const TX* __restrict__ xi_b = reinterpret_cast<const TX*>(a.x) + (int64_t)ci * a.x_sc;
    const TX* __restrict__ xj_b = reinterpret_cast<const TX*>(a.x) + (int64_t)cj * a.x_sc;
    const float* __restrict__ gi_b = reinterpret_cast<const float*>(a.g) + (int64_t)ci * a.g_sc;
    const float* __restrict__ gj_b = reinterpret_cast<const float*>(a.g) + (int64_t)cj * a.g_sc;

    const int64_t lo = a.units_total * blockIdx.x / gridDim.x;
    const int64_t hi = a.units_total * (blockIdx.x + 1) / gridDim.x;
    UnitCursor w;
    w.init(lo + rtid, a);

    // per-thread staging slots: [stage][plane 0..3 = x_i, x_j, g_i, g_j][thread] x 16 B, filled by cp.async
    // one iteration ahead, so the global-load latency is off the critical path without costing registers
    constexpr int kXB = sizeof(TX) * 4;  // bytes of 4 elements of x
    constexpr uint32_t kPlaneStride = kPThreads * 16;
    const uint32_t my_stage = smem_u32(stage_smem) + threadIdx.x * 16;
#define ECO_SLOT(st, plane) (my_stage + (uint32_t)((st) * 4 + (plane)) * kPlaneStride)
#define ECO_ISSUE(st)                                              \
    do {                                                           \
        cp_async_s<kXB, false>(ECO_SLOT(st, 0), xi_b + w.xoff);    \
        cp_async_s<kXB, false>(ECO_SLOT(st, 1), xj_b + w.xoff);    \
        cp_async_s<16, false>(ECO_SLOT(st, 2), gi_b + w.goff);     \
        cp_async_s<16, false>(ECO_SLOT(st, 3), gj_b + w.goff);     \
    } while (0)

    f2 acc[kRAcc];
#pragma unroll
    for (int k = 0; k < kRAcc; ++k) acc[k] = splat(0.f);
    int since_flush = 0;
    bool any_nonbinary = false;
    const int iters = (int)((hi - lo + kRoleThreads - 1) / kRoleThreads);
    if (w.q < hi) ECO_ISSUE(0);
    cp_async_commit();

    for (int it = 0; it < iters; ++it) {
        const bool active = w.q < hi;
        w.advance(kRoleThreads);
        const int st = it & 1;
        if (w.q < hi) {
            if (st) ECO_ISSUE(0);
            else ECO_ISSUE(1);
        }
        cp_async_commit();
        cp_async_wait<1>();  // this iteration's copies (committed one iteration ago) have landed
        if (active) {
            const uint32_t sbase = my_stage + (uint32_t)st * 4 * kPlaneStride;
            const float4 xi4 = lds_unit<TX>(sbase), xj4 = lds_unit<TX>(sbase + kPlaneStride);
            const float4 gi4 = lds_f4(sbase + 2 * kPlaneStride), gj4 = lds_f4(sbase + 3 * kPlaneStride);
            f2 pi[2] = {make_float2(xi4.x, xi4.y), make_float2(xi4.z, xi4.w)};
            f2 pj[2] = {make_float2(xj4.x, xj4.y), make_float2(xj4.z, xj4.w)};
            const f2 gi[2] = {make_float2(gi4.x, gi4.y), make_float2(gi4.z, gi4.w)};
            const f2 gj[2] = {make_float2(gj4.x, gj4.y), make_float2(gj4.z, gj4.w)};
            f2 d[2];
            if (LOGITS) {
                const f2 zi[2] = {pi[0], pi[1]}, zj[2] = {pj[0], pj[1]};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    pi[h] = sigmoid_fast2(zi[h]);
                    pj[h] = sigmoid_fast2(zj[h]);
                    d[h] = abs2(fma2(pj[h], splat(-1.0f), pi[h]));
                }
                if (fminf(fminf(d[0].x, d[0].y), fminf(d[1].x, d[1].y)) < kTieEps) {
                    // rare: the sign of (p_i - p_j) must come from ATen's exact sigmoid bits
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (d[h].x < kTieEps) { pi[h].x = sigmoid_exact(zi[h].x); pj[h].x = sigmoid_exact(zj[h].x); d[h].x = fabsf(pi[h].x - pj[h].x); }
                        if (d[h].y < kTieEps) { pi[h].y = sigmoid_exact(zi[h].y); pj[h].y = sigmoid_exact(zj[h].y); d[h].y = fabsf(pi[h].y - pj[h].y); }
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) d[h] = abs2(fma2(pj[h], splat(-1.0f), pi[h]));
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) role_pair_stats<LOGITS>(pi[h], pj[h], gi[h], gj[h], d[h], acc);
            if (chan_is_i) {  // warp-uniform: which of the two channels' plain leaf this role owns
#pragma unroll
                for (int h = 0; h < 2; ++h) role_plain_stats<LOGITS>(pi[h], gi[h], acc);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) role_plain_stats<LOGITS>(pj[h], gj[h], acc);
            }
        }
        if (++since_flush == kPFlushIters) {
            any_nonbinary |= flush_role_acc(acc, sm.warp_slots[warp], lane);
            since_flush = 0;
        }
    }
    cp_async_wait<0>();
#undef ECO_ISSUE
#undef ECO_SLOT
    any_nonbinary |= flush_role_acc(acc, sm.warp_slots[warp], lane);
    if (__syncthreads_or(any_nonbinary)) {
        // rare slow pass: some label is not exactly 0 or 1 -> exact transcendental corrections for this CTA's range
        const float* gi_s = reinterpret_cast<const float*>(a.g) + (int64_t)ci * a.g_sc;
        const float* gj_s = reinterpret_cast<const float*>(a.g) + (int64_t)cj * a.g_sc;
        for (int64_t q = lo + rtid; q < hi; q += kRoleThreads) {
            const int64_t n = q / a.units_per_plane, off = (q - n * a.units_per_plane) * 4;
            for (int e = 0; e < 4; ++e)
                label_corrections_role(gi_s[n * a.g_sn + off + e], gj_s[n * a.g_sn + off + e], role, sm.corr);
        }
        __syncthreads();
    }
    if (threadIdx.x < 96) {
        const int r = threadIdx.x >> 5, k = threadIdx.x & 31;
        double v = 0.0;
#pragma unroll
        for (int wq = 0; wq < kRoleWarps; ++wq) v += sm.warp_slots[r * kRoleWarps + wq][k];
        sm.role_sums[r][k] = v;
    }
    __syncthreads();
    double* mine = partials + (int64_t)blockIdx.x * kNAcc;
    if (threadIdx.x < kNAcc)
        mine[threadIdx.x] = packed_to_layout<LOGITS>(sm.role_sums, sm.corr, threadIdx.x, (double)(hi - lo) * 4.0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        sm.is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (sm.is_last) {
        __threadfence();
        // 100 sums x gridDim partials: 3 threads per sum, each a strided third, fixed order -> deterministic
        double* scratch = &sm.warp_slots[0][0];  // 384 doubles
        const int idx = threadIdx.x / 3, part = threadIdx.x % 3;
        double v = 0.0;
        if (idx < kNAcc)
            for (int i = part; i < (int)gridDim.x; i += 3) v += __ldcg(partials + (int64_t)i * kNAcc + idx);
        __syncthreads();
        scratch[threadIdx.x] = v;
        __syncthreads();
        double total = 0.0;
        if (threadIdx.x < kNAcc) total = scratch[3 * threadIdx.x] + scratch[3 * threadIdx.x + 1] + scratch[3 * threadIdx.x + 2];
        if (threadIdx.x < kNAcc) acc_out[threadIdx.x] = total;
        if (threadIdx.x == 0) *counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------------------------
struct PCoef {
    f2 u[12][6];  // plain c (0..2), then 3 + 3p + k for U_k of pair p: {sb', sab, sbb2, sp, fl, -}
    f2 i[9][2];   // 3p + k for I_k of pair p: {sa, sab}
};

// fill from the 21 LeafCoef of the leaf order (3 channel leaves, then per pair I1,U1,I2,U2,I3,U3)
template <bool UNIT>
__device__ __forceinline__ void fill_pcoef(PCoef& pc, const LeafCoef* cf, int t) {
    if (t >= ECO_C3_NLEAF) return;
    const LeafCoef c = cf[t];
    int ul = -1, il = -1;
    if (t < 3) ul = t;
    else {
        const int p = (t - 3) / 6, k = (t - 3) % 6;
        if (k & 1) ul = 3 + 3 * p + (k >> 1);
        else il = 3 * p + (k >> 1);
    }
    if (ul >= 0) {
        pc.u[ul][0] = splat(UNIT ? c.sb + 0.5f * c.sp : c.sb);
        pc.u[ul][1] = splat(c.sab);
        pc.u[ul][2] = splat(c.sbb2);
        pc.u[ul][3] = splat(c.sp);
        pc.u[ul][4] = splat(c.fl);
        pc.u[ul][5] = splat(0.f);
    } else {
        pc.i[il][0] = splat(c.sa);
        pc.i[il][1] = splat(c.sab);
    }
}

// d T / d b of a leaf with a = label, b in slot 2 (plain and union leaves)
template <bool UNIT, bool SIG, bool FL>
__device__ __forceinline__ f2 leaf_gb2(const f2 (&c)[6], f2 a, f2 b) {
    f2 r;
    if (!SIG) {
        r = fma2(c[1], a, fma2(c[2], b, c[0]));
    } else if (UNIT) {
        const f2 t = mul2(b, b);
        f2 s = fma2(splat(kSgS3), t, splat(kSgS2));
        s = fma2(s, t, splat(kSgS1));
        s = fma2(s, t, splat(kSgS0));
        const f2 w = fma2(c[3], s, c[2]);
        r = fma2(c[1], a, fma2(b, w, c[0]));
    } else {
        const f2 sg = make_float2(sigmoid_fast(b.x), sigmoid_fast(b.y));
        r = fma2(c[3], sg, fma2(c[1], a, fma2(c[2], b, c[0])));
    }
    if (FL) {  // + c_fl * d/db[-(1-b)^1.5 log(b+eps)]
        const f2 om = fma2(b, splat(-1.0f), splat(1.0f));
        const f2 s = make_float2(sqrt_approx(om.x), sqrt_approx(om.y));
        const f2 be = add2(b, splat(kEps));
        const f2 l = make_float2(lg2_approx(be.x), lg2_approx(be.y));
        const f2 rc = make_float2(rcp_approx(be.x), rcp_approx(be.y));
        const f2 v = fma2(mul2(s, splat(1.5f * kLn2)), l, mul2(mul2(om, s), mul2(rc, splat(-1.0f))));
        r = fma2(c[4], v, r);
    }
    return r;
}

// -sign(v) with sign(0) = 0 (torch.abs backward), as +-1.0f / 0.0f
__device__ __forceinline__ float neg_sign0(float v) {
    const float s = __uint_as_float((__float_as_uint(v) & 0x80000000u) ^ 0xbf800000u);  // -copysign(1, v)
    return v == 0.f ? 0.f : s;
}

template <bool UNIT, bool SIG, bool FL>
__device__ __forceinline__ void pixel_pair_grad(const f2 (&x)[3], const f2 (&g)[3], const f2 (&diffs)[3],
                                                const PCoef& pc, f2 (&gx)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gx[c] = leaf_gb2<UNIT, SIG, FL>(pc.u[c], g[c], x[c]);
    f2 hh[2];
    hh[0] = fma2(x[0], splat(-0.5f), splat(0.5f));
    hh[1] = fma2(x[1], splat(-0.5f), splat(0.5f));
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        const f2 xi = x[i], xj = x[j], gi = g[i], gj = g[j], h = hh[i];
        const f2 diff = diffs[p];
        const f2 d = abs2(diff);
        const f2 ns = make_float2(neg_sign0(diff.x), neg_sign0(diff.y));
        const f2 gd = abs2(fma2(gj, splat(-1.0f), gi));
        const f2 q = mul2(d, xi);
        const f2 nxis = mul2(xi, ns);                    // -x_i s   = d q / d x_j
        const f2 dq = fma2(nxis, splat(-1.0f), d);       // d + x_i s = d q / d x_i
        f2 gi_acc, gj_acc;
        {   // I1: a = x_i x_j, b = g_j
            const f2 ga = fma2(pc.i[3 * p + 0][1], gj, pc.i[3 * p + 0][0]);
            gi_acc = mul2(ga, xj);
            gj_acc = mul2(ga, xi);
        }
        {   // U1: b = u(x_i, x_j) = x_i + x_j h
            const f2 gb = leaf_gb2<UNIT, SIG, FL>(pc.u[3 + 3 * p + 0], gi, fma2(xj, h, xi));
            gi_acc = fma2(gb, fma2(xj, splat(-0.5f), splat(1.0f)), gi_acc);
            gj_acc = fma2(gb, h, gj_acc);
        }
        {   // I2: a = x_i d, b = gd
            const f2 ga = fma2(pc.i[3 * p + 1][1], gd, pc.i[3 * p + 1][0]);
            gi_acc = fma2(ga, dq, gi_acc);
            gj_acc = fma2(ga, nxis, gj_acc);
        }
        {   // U2: b = u(x_i, d): du/dx_i = 1 - d/2 + h s, du/dx_j = -h s
            const f2 gb = leaf_gb2<UNIT, SIG, FL>(pc.u[3 + 3 * p + 1], gi, fma2(d, h, xi));
            const f2 nhs = mul2(h, ns);
            const f2 dui = fma2(nhs, splat(-1.0f), fma2(d, splat(-0.5f), splat(1.0f)));
            gi_acc = fma2(gb, dui, gi_acc);
            gj_acc = fma2(gb, nhs, gj_acc);
        }
        {   // I3: a = x_i^2 d: da/dx_i = x_i (d + dq), da/dx_j = x_i (-x_i s)
            const f2 ga = fma2(pc.i[3 * p + 2][1], gd, pc.i[3 * p + 2][0]);
            gi_acc = fma2(ga, mul2(xi, add2(d, dq)), gi_acc);
            gj_acc = fma2(ga, mul2(xi, nxis), gj_acc);
        }
        {   // U3: b = u(x_i, q): du/dx_i = 1 - q/2 + h dq, du/dx_j = h (-x_i s)
            const f2 gb = leaf_gb2<UNIT, SIG, FL>(pc.u[3 + 3 * p + 2], gi, fma2(q, h, xi));
            const f2 dui = fma2(h, dq, fma2(q, splat(-0.5f), splat(1.0f)));
            gi_acc = fma2(gb, dui, gi_acc);
            gj_acc = fma2(gb, mul2(h, nxis), gj_acc);
        }
        gx[i] = add2(gx[i], gi_acc);
        gx[j] = add2(gx[j], gj_acc);
    }
}

template <typename TX, bool LOGITS, bool SIG, bool FL, int THREADS>
__device__ __forceinline__ void grad_phase_packed(const CompGradArgs& ga, const PCoef& pc, char* stage_smem, bool reverse) {
    const CompArgs& a = ga.a;
    const TX* __restrict__ xb = reinterpret_cast<const TX*>(a.x);
    const float* __restrict__ gb = reinterpret_cast<const float*>(a.g);
    TX* __restrict__ ob = reinterpret_cast<TX*>(ga.gx);
    const int64_t lo = a.units_total * blockIdx.x / gridDim.x;
    const int64_t hi = a.units_total * (blockIdx.x + 1) / gridDim.x;
    const int iters = (int)((hi - lo + THREADS - 1) / THREADS);
    constexpr int kXB = sizeof(TX) * 4;
    constexpr uint32_t kPlaneStride = THREADS * 16;
    const uint32_t my_stage = smem_u32(stage_smem) + threadIdx.x * 16;
    // cursor of the unit being STAGED (one iteration ahead of the one being processed); pass 2 may walk backwards
    UnitCursor w;
    w.init(lo + (int64_t)(reverse ? (iters - 1) : 0) * THREADS + threadIdx.x, a);
#define ECO_ISSUE2(st)                                                                                     \
    do {                                                                                                   \
        if (w.q >= lo && w.q < hi) {                                                                       \
            _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                                \
                cp_async_s<kXB, true>(my_stage + (uint32_t)((st) * 6 + c) * kPlaneStride, xb + w.xoff + c * a.x_sc);     \
                cp_async_s<16, true>(my_stage + (uint32_t)((st) * 6 + 3 + c) * kPlaneStride, gb + w.goff + c * a.g_sc); \
            }                                                                                              \
        }                                                                                                  \
        cp_async_commit();                                                                                 \
    } while (0)
    ECO_ISSUE2(0);
    for (int it = 0; it < iters; ++it) {
        const int64_t q = w.q;            // unit processed in this iteration
        const int64_t xoff = w.xoff;      // = n * x_sn + off * 4; the gradient is written with gx_sn, see below
        const int64_t n = w.n, off4 = w.off * 4;
        if (reverse) w.retreat(THREADS);
        else w.advance(THREADS);
        const int st = it & 1;
        if (it + 1 < iters) {
            if (st) ECO_ISSUE2(0);
            else ECO_ISSUE2(1);
        } else {
            cp_async_commit();
        }
        cp_async_wait<1>();
        (void)xoff;
        if (q >= hi) continue;
        const int64_t off = off4;
        float xv[3][4], gv[3][4], ov[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 xq = lds_unit<TX>(my_stage + (uint32_t)(st * 6 + c) * kPlaneStride);
            const float4 gq = lds_f4(my_stage + (uint32_t)(st * 6 + 3 + c) * kPlaneStride);
            xv[c][0] = xq.x; xv[c][1] = xq.y; xv[c][2] = xq.z; xv[c][3] = xq.w;
            gv[c][0] = gq.x; gv[c][1] = gq.y; gv[c][2] = gq.z; gv[c][3] = gq.w;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            f2 x[3], g[3], gx[3], diffs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                x[c] = make_float2(xv[c][2 * h], xv[c][2 * h + 1]);
                g[c] = make_float2(gv[c][2 * h], gv[c][2 * h + 1]);
            }
            if (LOGITS) {
                const f2 z0 = x[0], z1 = x[1], z2 = x[2];
                x[0] = sigmoid_fast2(z0);
                x[1] = sigmoid_fast2(z1);
                x[2] = sigmoid_fast2(z2);
#pragma unroll
                for (int p = 0; p < 3; ++p) diffs[p] = fma2(x[pair_j(p)], splat(-1.0f), x[pair_i(p)]);
                const float dx = fminf(fminf(fabsf(diffs[0].x), fabsf(diffs[1].x)), fabsf(diffs[2].x));
                const float dy = fminf(fminf(fabsf(diffs[0].y), fabsf(diffs[1].y)), fabsf(diffs[2].y));
                if (fminf(dx, dy) < kTieEps) {  // rare: exact ATen sigmoid bits decide the sign of the kink
                    if (dx < kTieEps) { x[0].x = sigmoid_exact(z0.x); x[1].x = sigmoid_exact(z1.x); x[2].x = sigmoid_exact(z2.x); }
                    if (dy < kTieEps) { x[0].y = sigmoid_exact(z0.y); x[1].y = sigmoid_exact(z1.y); x[2].y = sigmoid_exact(z2.y); }
#pragma unroll
                    for (int p = 0; p < 3; ++p) diffs[p] = fma2(x[pair_j(p)], splat(-1.0f), x[pair_i(p)]);
                }
            } else {
#pragma unroll
                for (int p = 0; p < 3; ++p) diffs[p] = fma2(x[pair_j(p)], splat(-1.0f), x[pair_i(p)]);
            }
            pixel_pair_grad<LOGITS, SIG, FL>(x, g, diffs, pc, gx);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                f2 o = gx[c];
                if (LOGITS) o = mul2(o, mul2(x[c], fma2(x[c], splat(-1.0f), splat(1.0f))));
                ov[c][2 * h] = o.x;
                ov[c][2 * h + 1] = o.y;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) Vec4<TX>::store(ob + n * ga.gx_sn + c * ga.gx_sc + off, ov[c]);
    }
    cp_async_wait<0>();
#undef ECO_ISSUE2
}

}  // namespace eco
