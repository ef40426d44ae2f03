// Fused 3-organ composite loss (whole_body > ventral+dorsal > dorsal): 21 (a,b) leaves per pixel
// from ONE read of x[:,0..2] and g[:,0..2].  Replaces ess/loss_composite.py:21-94 with
// composite_set_theory=True, C == 3 (see include/ecoloss.h).
//
// Pass 1 (stats): per pixel, form p_c (= sigmoid(z_c) when from_logits), the pair operands
//   d=|p_i-p_j|, m1=p_i p_j, m2=p_i d, m3=p_i (d p_i), u1=u(p_i,p_j), u2=u(p_i,d), u3=u(p_i,d p_i)
//   and accumulate 84 fp32 sums per thread; every kFlushIters iterations the warp folds them
//   (transposed butterfly, 31 shuffles per 32 sums) into per-warp fp64 slots in shared memory.
//   Labels that are exactly 0/1 let the label-only transcendental sums collapse to counts; any other
//   label value takes a rare slow path that accumulates exact corrections, so the result is right for
//   arbitrary labels.
// Pass 2 (grad): chain rule through the operands with 21 x 6 global coefficients.
#include <cooperative_groups.h>
#include <string.h>

#include "eco_common.cuh"

namespace eco {

constexpr int kCThreads = 256;
constexpr int kCWarps = kCThreads / 32;
constexpr int kNAcc = 100;      // == ECO_C3_NACC
constexpr int kNThreadAcc = 84; // indices 1..84 live in registers
constexpr int kFlushIters = 8;
constexpr int kMaxCompCtas = 148 * 4;

static_assert(kNAcc == ECO_C3_NACC, "layout mismatch with ecoloss.h");

// ---- accumulator layout ---------------------------------------------------------------------
// 0            n (pixels)
// 1..3         G_c   = sum g_c
// 4..6         GD_p  = sum |g_i-g_j|           p = 0:(0,1) 1:(0,2) 2:(1,2)
// 7+5c+k       channel leaf c: k = 0 X (sum x) 1 XX 2 GX 3 SPX 4 FLX(log2 units until flushed)
// 22+21p+k     pair p: 0 M1 1 M1G 2 U1 3 UU1 4 GU1 5 SPU1 6 FLU1 | 7 M2 8 M2G 9 U2 10 UU2 11 GU2 12 SPU2 13 FLU2
//                      | 14 M3 15 M3G 16 U3 17 UU3 18 GU3 19 SPU3 20 FLU3
// 85+3L+k      label-b corrections (non-binary labels only), L = 0:g1 1:g2 2:gd01 3:gd02 4:gd12,
//              k = 0 sum(b*b-b)  1 sum(SP(b)-lin)  2 sum(FL(b)-lin)
constexpr int A_N = 0, A_G = 1, A_GD = 4, A_CH = 7, A_PAIR = 22, A_CORR = 85;

__host__ __device__ constexpr int pair_i(int p) { return p == 2 ? 1 : 0; }
__host__ __device__ constexpr int pair_j(int p) { return p == 0 ? 1 : 2; }

// natural-log constants of the two label values
constexpr double kSP0 = 0.6931471805599453094;       // softplus term at b = 0: log 2
constexpr double kSP1 = 1.3132616875182228340;       // at b = 1: 1 + log(1 + e^-1)
constexpr double kFL0 = 16.118095533458650;          // -(1-0)^1.5 log(0 + fp32(1e-7))

// reference op order of ess/loss_composite.py:94, no FMA contraction across the ops
__device__ __forceinline__ float union_operand(float sp, float p) {
    return __fadd_rn(__fmul_rn(sp, __fsub_rn(1.0f, p)), __fmul_rn(__fadd_rn(__fmul_rn(sp, p), p), 0.5f));
}

struct CompArgs {
    const void* x;
    const void* g;
    int64_t x_sn, x_sc, g_sn, g_sc;
    int32_t N;
    int64_t HW;
    int64_t units_per_plane;  // HW / VEC
    int64_t units_total;      // N * units_per_plane
};

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 4) Vec4<T>::load(p, v);
    else v[0] = Vec4<T>::load1(p);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) Vec4<T>::store(p, v);
    else Vec4<T>::store1(p, v[0]);
}

// thread accumulators: acc[k] <-> layout index k+1
template <bool UNIT>
__device__ __forceinline__ void leaf_b_terms(float b, float& sp_acc, float& fl_acc) {
    // SP in natural units; FL in log2 units with the sign folded (fl_acc accumulates (1-b)^1.5 log2(b+eps))
    if (UNIT) sp_acc += fmaf(softplus_neg_abs_log2(b), kLn2, b);
    else sp_acc += fmaf(softplus_neg_abs_log2(b), kLn2, fmaxf(b, 0.f));
    fl_acc += focal_fg_log2(b);
}

template <bool UNIT>
__device__ __forceinline__ void pixel_stats(const float (&x)[3], const float (&g)[3], float (&acc)[kNThreadAcc]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        acc[A_G - 1 + c] += g[c];
        float* ch = &acc[A_CH - 1 + 5 * c];
        ch[0] += x[c];
        ch[1] = fmaf(x[c], x[c], ch[1]);
        ch[2] = fmaf(g[c], x[c], ch[2]);
        leaf_b_terms<UNIT>(x[c], ch[3], ch[4]);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        const float xi = x[i], xj = x[j], gi = g[i], gj = g[j];
        const float d = fabsf(xi - xj);
        const float gd = fabsf(gi - gj);
        acc[A_GD - 1 + p] += gd;
        float* pa = &acc[A_PAIR - 1 + 21 * p];
        const float m1 = xi * xj;
        const float m2 = xi * d;
        const float q = d * xi;
        const float m3 = xi * q;
        const float u1 = union_operand(xi, xj);
        const float u2 = union_operand(xi, d);
        const float u3 = union_operand(xi, q);
        pa[0] += m1;  pa[1] = fmaf(m1, gj, pa[1]);
        pa[2] += u1;  pa[3] = fmaf(u1, u1, pa[3]);  pa[4] = fmaf(gi, u1, pa[4]);  leaf_b_terms<UNIT>(u1, pa[5], pa[6]);
        pa[7] += m2;  pa[8] = fmaf(m2, gd, pa[8]);
        pa[9] += u2;  pa[10] = fmaf(u2, u2, pa[10]); pa[11] = fmaf(gi, u2, pa[11]); leaf_b_terms<UNIT>(u2, pa[12], pa[13]);
        pa[14] += m3; pa[15] = fmaf(m3, gd, pa[15]);
        pa[16] += u3; pa[17] = fmaf(u3, u3, pa[17]); pa[18] = fmaf(gi, u3, pa[18]); leaf_b_terms<UNIT>(u3, pa[19], pa[20]);
    }
}

// rare path: exact corrections for label values other than 0/1 (double math, shared atomics)
__device__ __noinline__ void label_corrections(const float (&g)[3], double* corr /* smem [15] */) {
    const float lb[5] = {g[1], g[2], fabsf(g[0] - g[1]), fabsf(g[0] - g[2]), fabsf(g[1] - g[2])};
    for (int L = 0; L < 5; ++L) {
        const double b = (double)lb[L];
        if (b == 0.0 || b == 1.0) continue;
        const double be = (double)(lb[L] + kEps);  // fp32 add like the reference
        const double sp = fmax(b, 0.0) + log1p(exp(-fabs(b)));
        const double fl = -pow(1.0 - b, 1.5) * log(be);
        atomicAdd(&corr[3 * L + 0], b * b - b);
        atomicAdd(&corr[3 * L + 1], sp - ((1.0 - b) * kSP0 + b * kSP1));
        atomicAdd(&corr[3 * L + 2], fl - (1.0 - b) * kFL0);
    }
}

// fold 32 per-lane values so that lane l ends up with the warp total of v[l]
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = lane & s;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

__device__ __forceinline__ void flush_thread_acc(float (&acc)[kNThreadAcc], double* warp_slot /* smem [96] */, int lane) {
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int k = grp * 32 + i;
            v[i] = (k < kNThreadAcc) ? acc[k] : 0.f;
        }
        const float tot = butterfly32(v, lane);
        warp_slot[grp * 32 + lane] += (double)tot;
    }
#pragma unroll
    for (int k = 0; k < kNThreadAcc; ++k) acc[k] = 0.f;
}

// is layout index `idx` accumulated in log2 units with the sign folded?  (FLX, FLU*)
__host__ __device__ inline bool is_focal_slot(int idx) {
    if (idx >= A_CH && idx < A_PAIR) return (idx - A_CH) % 5 == 4;
    if (idx >= A_PAIR && idx < A_CORR) {
        const int k = (idx - A_PAIR) % 21;
        return k == 6 || k == 13 || k == 20;
    }
    return false;
}

struct StatsSmem {
    double warp_slots[kCWarps][96];
    double corr[15];
    bool is_last;
};

// Phase 1 body, shared by the stand-alone stats kernel and the fused cooperative kernel.  On return the
// LAST CTA to arrive has written acc_out[0..100) (visible device-wide after a grid barrier / kernel end).
template <typename TX, int VEC, bool LOGITS>
__device__ __forceinline__ void stats_phase(const CompArgs& a, StatsSmem& sm, unsigned int* __restrict__ counter,
                                            double* __restrict__ partials, double* __restrict__ acc_out) {
    double (*warp_slots)[96] = sm.warp_slots;
    double* corr = sm.corr;
    bool& is_last = sm.is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kCWarps * 96; i += kCThreads) (&warp_slots[0][0])[i] = 0.0;
    if (threadIdx.x < 15) corr[threadIdx.x] = 0.0;
    __syncthreads();

    const TX* __restrict__ xb = reinterpret_cast<const TX*>(a.x);
    const float* __restrict__ gb = reinterpret_cast<const float*>(a.g);

    const int64_t lo = a.units_total * blockIdx.x / gridDim.x;
    const int64_t hi = a.units_total * (blockIdx.x + 1) / gridDim.x;
    int64_t q = lo + threadIdx.x;
    int64_t n = q / a.units_per_plane;
    int64_t off = q - n * a.units_per_plane;

    float acc[kNThreadAcc];
#pragma unroll
    for (int k = 0; k < kNThreadAcc; ++k) acc[k] = 0.f;
    int since_flush = 0;

    // warp-uniform trip count: the in-loop flush shuffles with a full mask
    const int iters = (int)((hi - lo + kCThreads - 1) / kCThreads);
    for (int it = 0; it < iters; ++it, q += kCThreads) {
        if (q < hi) {
            float xv[3][VEC], gv[3][VEC];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                load_vec<TX, VEC>(xb + n * a.x_sn + c * a.x_sc + off * VEC, xv[c]);
                load_vec<float, VEC>(gb + n * a.g_sn + c * a.g_sc + off * VEC, gv[c]);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float x[3], g[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    x[c] = LOGITS ? sigmoid_exact(xv[c][v]) : xv[c][v];
                    g[c] = gv[c][v];
                }
                pixel_stats<LOGITS>(x, g, acc);
                const bool nonbin = (g[0] != 0.f && g[0] != 1.f) || (g[1] != 0.f && g[1] != 1.f) || (g[2] != 0.f && g[2] != 1.f);
                if (nonbin) label_corrections(g, corr);
            }
            off += kCThreads;
            while (off >= a.units_per_plane) {
                off -= a.units_per_plane;
                ++n;
            }
        }
        if (++since_flush == kFlushIters) {
            flush_thread_acc(acc, warp_slots[warp], lane);
            since_flush = 0;
        }
    }
    // every lane of a warp must take part in the shuffles: flush unconditionally
    flush_thread_acc(acc, warp_slots[warp], lane);
    __syncthreads();

    double* mine = partials + (int64_t)blockIdx.x * kNAcc;
    for (int idx = threadIdx.x; idx < kNAcc; idx += kCThreads) {
        double v = 0.0;
        if (idx == A_N) {
            v = (double)(hi - lo) * VEC;
        } else if (idx < A_CORR) {
#pragma unroll
            for (int w = 0; w < kCWarps; ++w) v += warp_slots[w][idx - 1];
            if (is_focal_slot(idx)) v *= -kLn2d;
        } else {
            v = corr[idx - A_CORR];
        }
        mine[idx] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int idx = warp; idx < kNAcc; idx += kCWarps) {
            double v = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(partials + (int64_t)i * kNAcc + idx);
            v = warp_sum(v);
            if (lane == 0) acc_out[idx] = v;
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}

template <typename TX, int VEC, bool LOGITS>
__global__ void __launch_bounds__(kCThreads, 1)
composite3_stats_kernel(CompArgs a, unsigned int* __restrict__ counter, double* __restrict__ partials,
                        double* __restrict__ acc_out) {
    __shared__ StatsSmem sm;
    stats_phase<TX, VEC, LOGITS>(a, sm, counter, partials, acc_out);
}

// ---------------------------------------------------------------------------------------------
// finalize: 100 sums -> 21 leaves' stats -> losses[7], jac[21][7][7]
// ---------------------------------------------------------------------------------------------
template <bool CG>
__device__ inline void composite_leaf_sums_t(const double* A_, int leaf, double* s /*[8]*/) {
    struct Rd { const double* p; __device__ double operator[](int i) const { return CG ? __ldcg(p + i) : p[i]; } };
    const Rd A{A_};
    const double n = A[A_N];
    s[S_N] = n;
    s[S_FLB] = 0.0;
    if (leaf < 3) {
        const Rd ch{A_ + A_CH + 5 * leaf};
        s[S_A] = A[A_G + leaf]; s[S_B] = ch[0]; s[S_BB] = ch[1]; s[S_AB] = ch[2]; s[S_SP] = ch[3]; s[S_FL] = ch[4];
        return;
    }
    const int p = (leaf - 3) / 6, t = (leaf - 3) % 6;
    const int i = pair_i(p), j = pair_j(p);
    const double* pa = A_ + A_PAIR + 21 * p;
    if (t & 1) {  // U-leaf: a = g_i, b = u_k
        const Rd u{pa + 2 + 7 * (t >> 1)};
        s[S_A] = A[A_G + i]; s[S_B] = u[0]; s[S_BB] = u[1]; s[S_AB] = u[2]; s[S_SP] = u[3]; s[S_FL] = u[4];
    } else {  // I-leaf: a = m_k, b = label (g_j for I1, gd for I2/I3)
        const Rd m{pa + 7 * (t >> 1)};
        const int L = (t == 0) ? (j - 1) : (2 + p);
        const double sb = (t == 0) ? A[A_G + j] : A[A_GD + p];
        const Rd co{A_ + A_CORR + 3 * L};
        s[S_A] = m[0]; s[S_AB] = m[1]; s[S_B] = sb;
        s[S_BB] = sb + co[0];
        s[S_SP] = (n - sb) * kSP0 + sb * kSP1 + co[1];
        s[S_FL] = (n - sb) * kFL0 + co[2];
    }
}

__device__ inline void composite_leaf_sums(const double* A, int leaf, double* s) { composite_leaf_sums_t<false>(A, leaf, s); }
__device__ inline void composite_leaf_sums_ldcg(const double* A, int leaf, double* s) { composite_leaf_sums_t<true>(A, leaf, s); }

struct CompFinArgs {
    double scale[ECO_C3_NLEAF];
};

__global__ void composite3_finalize_kernel(const double* __restrict__ A, CompFinArgs fa, const double* scale_dev,
                                           float* __restrict__ losses_out, double* __restrict__ jac_out,
                                           double* __restrict__ leaf_sums_out) {
    __shared__ double sl[ECO_C3_NLEAF][ECO_NLOSS];
    if (threadIdx.x < ECO_C3_NLEAF * ECO_NLOSS) {
        const int leaf = threadIdx.x / ECO_NLOSS, k = threadIdx.x % ECO_NLOSS;
        double s[ECO_NSTAT], jrow[ECO_NJAC];
        composite_leaf_sums(A, leaf, s);
        leaf_closed_form_row(s, 0.0, scale_dev ? scale_dev[leaf] : fa.scale[leaf], k, sl[leaf][k], jrow);
        if (jac_out)
            for (int j = 0; j < ECO_NJAC; ++j) jac_out[(leaf * ECO_NLOSS + k) * ECO_NJAC + j] = jrow[j];
        if (leaf_sums_out && k == 0)
            for (int j = 0; j < ECO_NSTAT; ++j) leaf_sums_out[leaf * ECO_NSTAT + j] = s[j];
    }
    __syncthreads();
    if (losses_out && threadIdx.x < ECO_NLOSS) {
        double v = 0.0;
        for (int l = 0; l < ECO_C3_NLEAF; ++l) v += sl[l][threadIdx.x];
        losses_out[threadIdx.x] = (float)v;
    }
}

// ---------------------------------------------------------------------------------------------
// gradient
// ---------------------------------------------------------------------------------------------
struct CompGradArgs {
    CompArgs a;
    void* gx;
    int64_t gx_sn, gx_sc;
};

__device__ __forceinline__ float leaf_gb(const LeafCoef& c, float a, float b, bool need_sig, bool need_fl) {
    float r = fmaf(c.sab, a, fmaf(c.sbb2, b, c.sb));
    if (need_sig) r = fmaf(c.sp, sigmoid_fast(b), r);
    if (need_fl) r = fmaf(c.fl, dfocal_fg(b), r);
    return r;
}

__device__ __forceinline__ void pixel_grad(const float (&x)[3], const float (&g)[3], const LeafCoef* __restrict__ cf,
                                           bool need_sig, bool need_fl, float (&gx)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gx[c] = leaf_gb(cf[c], g[c], x[c], need_sig, need_fl);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        const LeafCoef* L = cf + 3 + 6 * p;
        const float xi = x[i], xj = x[j], gi = g[i], gj = g[j];
        const float diff = xi - xj;
        const float d = fabsf(diff);
        const float s = (diff > 0.f ? 1.f : 0.f) - (diff < 0.f ? 1.f : 0.f);
        const float gd = fabsf(gi - gj);
        const float q = d * xi;
        const float xis = xi * s;
        const float dq_i = d + xis;  // d q / d x_i (also d m2 / d x_i)
        const float hp = 0.5f * (1.0f - xi);  // d u / d p
        float gi_acc = 0.f, gj_acc = 0.f;
        // I1: a = xi*xj, b = gj
        {
            const float ga = fmaf(L[0].sab, gj, L[0].sa);
            gi_acc = fmaf(ga, xj, gi_acc);
            gj_acc = fmaf(ga, xi, gj_acc);
        }
        // U1: a = gi, b = u(xi, xj)
        {
            const float gb = leaf_gb(L[1], gi, union_operand(xi, xj), need_sig, need_fl);
            gi_acc = fmaf(gb, 1.0f - 0.5f * xj, gi_acc);
            gj_acc = fmaf(gb, hp, gj_acc);
        }
        // I2: a = xi*d, b = gd
        {
            const float ga = fmaf(L[2].sab, gd, L[2].sa);
            gi_acc = fmaf(ga, dq_i, gi_acc);
            gj_acc = fmaf(ga, -xis, gj_acc);
        }
        // U2: a = gi, b = u(xi, d)
        {
            const float gb = leaf_gb(L[3], gi, union_operand(xi, d), need_sig, need_fl);
            gi_acc = fmaf(gb, fmaf(hp, s, 1.0f - 0.5f * d), gi_acc);
            gj_acc = fmaf(gb, -hp * s, gj_acc);
        }
        // I3: a = xi*(d*xi), b = gd
        {
            const float ga = fmaf(L[4].sab, gd, L[4].sa);
            gi_acc = fmaf(ga, fmaf(2.0f * xi, d, xi * xis), gi_acc);
            gj_acc = fmaf(ga, -xi * xis, gj_acc);
        }
        // U3: a = gi, b = u(xi, q), q = d*xi
        {
            const float gb = leaf_gb(L[5], gi, union_operand(xi, q), need_sig, need_fl);
            gi_acc = fmaf(gb, fmaf(hp, dq_i, 1.0f - 0.5f * q), gi_acc);
            gj_acc = fmaf(gb, -hp * xis, gj_acc);
        }
        gx[i] += gi_acc;
        gx[j] += gj_acc;
    }
}

template <typename TX, int VEC, bool LOGITS>
__device__ __forceinline__ void grad_phase(const CompGradArgs& ga, const LeafCoef* __restrict__ cf, bool need_sig,
                                           bool need_fl) {
    const CompArgs& a = ga.a;
    const TX* __restrict__ xb = reinterpret_cast<const TX*>(a.x);
    const float* __restrict__ gb = reinterpret_cast<const float*>(a.g);
    TX* __restrict__ ob = reinterpret_cast<TX*>(ga.gx);

    const int64_t lo = a.units_total * blockIdx.x / gridDim.x;
    const int64_t hi = a.units_total * (blockIdx.x + 1) / gridDim.x;
    int64_t q = lo + threadIdx.x;
    int64_t n = q / a.units_per_plane;
    int64_t off = q - n * a.units_per_plane;
    const int nthreads = blockDim.x;

    for (; q < hi; q += nthreads) {
        float xv[3][VEC], gv[3][VEC], ov[3][VEC];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            load_vec<TX, VEC>(xb + n * a.x_sn + c * a.x_sc + off * VEC, xv[c]);
            load_vec<float, VEC>(gb + n * a.g_sn + c * a.g_sc + off * VEC, gv[c]);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float x[3], g[3], gx[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                x[c] = LOGITS ? sigmoid_exact(xv[c][v]) : xv[c][v];
                g[c] = gv[c][v];
            }
            pixel_grad(x, g, cf, need_sig, need_fl, gx);
#pragma unroll
            for (int c = 0; c < 3; ++c) ov[c][v] = LOGITS ? gx[c] * ((1.0f - x[c]) * x[c]) : gx[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) store_vec<TX, VEC>(ob + n * ga.gx_sn + c * ga.gx_sc + off * VEC, ov[c]);
        off += nthreads;
        while (off >= a.units_per_plane) {
            off -= a.units_per_plane;
            ++n;
        }
    }
}

template <typename TX, int VEC, bool LOGITS>
__global__ void __launch_bounds__(kCThreads, 2)
composite3_grad_kernel(CompGradArgs ga, const double* __restrict__ jac, const float* __restrict__ upstream) {
    __shared__ LeafCoef cf[ECO_C3_NLEAF];
    if (threadIdx.x < ECO_C3_NLEAF) cf[threadIdx.x] = make_coef(jac + threadIdx.x * ECO_NLOSS * ECO_NJAC, upstream);
    __syncthreads();
    grad_phase<TX, VEC, LOGITS>(ga, cf, upstream[1] != 0.f, upstream[2] != 0.f);
}

// ---------------------------------------------------------------------------------------------
// fused: stats -> grid barrier -> closed forms (redundantly per CTA) -> gradient, ONE cooperative launch
// ---------------------------------------------------------------------------------------------
template <typename TX, int VEC, bool LOGITS>
__global__ void __launch_bounds__(kCThreads, 1)
composite3_fused_kernel(CompGradArgs ga, const double* __restrict__ scale_dev, const float* __restrict__ upstream,
                        unsigned int* __restrict__ counter,
                        double* __restrict__ partials, double* __restrict__ acc_glob, float* __restrict__ losses_out) {
    __shared__ StatsSmem sm;
    __shared__ LeafCoef cf[ECO_C3_NLEAF];
    __shared__ double sl[ECO_C3_NLEAF][ECO_NLOSS];
    stats_phase<TX, VEC, LOGITS>(ga.a, sm, counter, partials, acc_glob);

    // grid barrier (all CTAs are co-resident: cooperative launch); also orders the last CTA's acc_glob writes
    __threadfence();
    cooperative_groups::this_grid().sync();

    const int leaf = threadIdx.x;
    if (leaf < ECO_C3_NLEAF) {
        // each of the 21 threads needs only a handful of the 100 sums; read straight from L2
        double s[ECO_NSTAT];
        composite_leaf_sums_ldcg(acc_glob, leaf, s);
        LeafOut o;
        leaf_closed_form(s, 0.0, scale_dev[leaf], o);
        cf[leaf] = make_coef(&o.jac[0][0], upstream);
        for (int k = 0; k < ECO_NLOSS; ++k) sl[leaf][k] = o.loss[k];
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < ECO_NLOSS) {
        double v = 0.0;
        for (int l = 0; l < ECO_C3_NLEAF; ++l) v += sl[l][threadIdx.x];
        losses_out[threadIdx.x] = (float)v;
    }
    grad_phase<TX, VEC, LOGITS>(ga, cf, upstream[1] != 0.f, upstream[2] != 0.f);
}

// rare path of the packed kernels: one role's share of the transcendental label corrections for one pixel
// (sum(b^2 - b) comes from the second-moment accumulators)
__device__ __noinline__ void label_corrections_role(float gi, float gj, int role, double* corr) {
    float lb[2];
    int L[2];
    int n = 0;
    if (role != 2) { lb[n] = gj; L[n] = role; ++n; }       // g1 (role 0) / g2 (role 1); role 2 shares g2
    lb[n] = fabsf(gi - gj); L[n] = 2 + role; ++n;           // gd01 / gd02 / gd12
    for (int k = 0; k < n; ++k) {
        const double b = (double)lb[k];
        if (b == 0.0 || b == 1.0) continue;
        const double be = (double)(lb[k] + kEps);
        const double sp = fmax(b, 0.0) + log1p(exp(-fabs(b)));
        const double fl = -pow(1.0 - b, 1.5) * log(be);
        atomicAdd(&corr[3 * L[k] + 1], sp - ((1.0 - b) * kSP0 + b * kSP1));
        atomicAdd(&corr[3 * L[k] + 2], fl - (1.0 - b) * kFL0);
    }
}

}  // namespace eco

#include "eco_composite_packed.cuh"
#include "eco_composite_v2.cuh"
#include "eco_composite_v3.cuh"
#include "eco_multiclass_v2.cuh"

namespace eco {

template <typename TX, bool LOGITS>
__global__ void __launch_bounds__(kPThreads, 1)
composite3_stats_packed_kernel(CompArgs a, unsigned int* __restrict__ counter, double* __restrict__ partials,
                               double* __restrict__ acc_out) {
    extern __shared__ __align__(16) char stage_smem[];
    __shared__ PStatsSmem sm;
    stats_phase_packed<TX, LOGITS>(a, sm, stage_smem, counter, partials, acc_out);
}

// shared tail of the packed gradient kernels: coefficients -> (packed | scalar focal fallback) gradient pass
template <typename TX, bool LOGITS, int THREADS>
__device__ __forceinline__ void grad_dispatch_packed(const CompGradArgs& ga, LeafCoef* cf, PCoef& pc, char* stage_smem,
                                                     const float* __restrict__ upstream, bool reverse) {
    const bool need_sig = upstream[1] != 0.f, need_fl = upstream[2] != 0.f;
    fill_pcoef<LOGITS>(pc, cf, threadIdx.x);
    __syncthreads();
    // block-uniform dispatch on which of the 7 outputs carry gradient (train_multiclass.py:145 weights them 0/1)
    if (need_fl) {
        if (need_sig) grad_phase_packed<TX, LOGITS, true, true, THREADS>(ga, pc, stage_smem, reverse);
        else grad_phase_packed<TX, LOGITS, false, true, THREADS>(ga, pc, stage_smem, reverse);
    } else {
        if (need_sig) grad_phase_packed<TX, LOGITS, true, false, THREADS>(ga, pc, stage_smem, reverse);
        else grad_phase_packed<TX, LOGITS, false, false, THREADS>(ga, pc, stage_smem, reverse);
    }
}

constexpr int kGThreads = 256;   // stand-alone pass 2: fewer, fatter threads (255 registers, no spills)
template <typename TX, bool LOGITS>
__global__ void __launch_bounds__(kGThreads, 1)
composite3_grad_packed_kernel(CompGradArgs ga, const double* __restrict__ jac, const float* __restrict__ upstream) {
    extern __shared__ __align__(16) char stage_smem[];
    __shared__ LeafCoef cf[ECO_C3_NLEAF];
    __shared__ PCoef pc;
    if (threadIdx.x < ECO_C3_NLEAF) cf[threadIdx.x] = make_coef(jac + threadIdx.x * ECO_NLOSS * ECO_NJAC, upstream);
    __syncthreads();
    grad_dispatch_packed<TX, LOGITS, kGThreads>(ga, cf, pc, stage_smem, upstream, false);
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static bool c_aligned(const void* ptr, int64_t sn, int64_t sc, int dtype, int64_t HW) {
    const int64_t esz = dtype == ECO_BF16 ? 2 : 4;
    return (reinterpret_cast<uintptr_t>(ptr) % (4 * esz) == 0) && (sn % 4 == 0) && (sc % 4 == 0) && (HW % 4 == 0);
}

static int fill_comp(CompArgs& a, const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int vec) {
    a.x = x->ptr; a.g = g->ptr;
    a.x_sn = x->sn; a.x_sc = x->sc; a.g_sn = g->sn; a.g_sc = g->sc;
    a.N = N; a.HW = HW;
    a.units_per_plane = HW / vec;
    a.units_total = a.units_per_plane * N;
    return 0;
}

static int check_comp(const EcoView* x, const EcoView* g, int32_t N, int64_t HW) {
    if (N <= 0 || HW <= 0) { set_error("empty input (N=%d HW=%lld)", N, (long long)HW); return -2; }
    if (!x || !g || !x->ptr || !g->ptr) { set_error("null input view"); return -1; }
    if (x->dtype != ECO_F32 && x->dtype != ECO_BF16) { set_error("x dtype must be f32 or bf16"); return -4; }
    if (g->dtype != ECO_F32 && g->dtype != ECO_U8) { set_error("composite3: labels must be f32 or u8"); return -4; }
    return 0;
}
static int need_f32_labels(const EcoView* g, const char* what) {
    if (g->dtype != ECO_F32) { set_error("%s takes f32 labels (byte labels: eco_composite3_step on fp32 logits with 16-byte aligned planes)", what); return -4; }
    return 0;
}

// opt in to > 48 KB of dynamic shared memory once per process and device (the attribute is per device)
template <typename K>
static int set_smem_attr(K kernel) {
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes), "cudaFuncSetAttribute(smem)");
}
static int ensure_packed_smem() {
    static thread_local int done_for_device[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -6;
    if (done_for_device[dev]) return 0;
    int rc = 0;
#define ECO_SMEM_ALL(TX)                                                                           \
    rc = rc ? rc : set_smem_attr(composite3_stats_packed_kernel<TX, true>);                        \
    rc = rc ? rc : set_smem_attr(composite3_stats_packed_kernel<TX, false>);                       \
    rc = rc ? rc : set_smem_attr(composite3_grad_packed_kernel<TX, true>);                         \
    rc = rc ? rc : set_smem_attr(composite3_grad_packed_kernel<TX, false>);
    ECO_SMEM_ALL(float)
    ECO_SMEM_ALL(__nv_bfloat16)
#undef ECO_SMEM_ALL
    rc = rc ? rc : check_cuda(cudaFuncSetAttribute(v2::multiclass3_fused_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes), "cudaFuncSetAttribute(smem, multiclass v2)");
    rc = rc ? rc : check_cuda(cudaFuncSetAttribute(v2::multiclass3_fused_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes), "cudaFuncSetAttribute(smem, multiclass v2 probs)");
    rc = rc ? rc : check_cuda(cudaFuncSetAttribute(v2::composite3_grad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes), "cudaFuncSetAttribute(smem, grad v2)");
#define ECO_V3_ATTR(TX, TG, PR) rc = rc ? rc : check_cuda(cudaFuncSetAttribute(v2::composite3_fused_v3_kernel<TX, TG, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::Stage3<TX, TG>::kSmem), "cudaFuncSetAttribute(smem, fused v3)")
    ECO_V3_ATTR(float, float, false); ECO_V3_ATTR(float, uint8_t, false); ECO_V3_ATTR(__nv_bfloat16, float, false); ECO_V3_ATTR(__nv_bfloat16, uint8_t, false);
    ECO_V3_ATTR(float, float, true); ECO_V3_ATTR(float, uint8_t, true);
#undef ECO_V3_ATTR
    if (!rc) done_for_device[dev] = 1;
    return rc;
}

static int comp_grid(int device, int64_t units, int ctas_per_sm, int threads_per_unit_stride) {
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -1;
    int64_t g = (int64_t)sms * ctas_per_sm;
    const int64_t need = (units + threads_per_unit_stride - 1) / threads_per_unit_stride;
    if (g > need) g = need;
    if (g > kMaxCompCtas) g = kMaxCompCtas;
    if (g < 1) g = 1;
    return (int)g;
}

// second-generation kernels: fp32 logits, 16-byte aligned planes, sums inside the fixed-point range of V2Ws
static bool v2_eligible(const EcoView* x, int32_t from_logits, int vec, int32_t N, int64_t HW) {
    return vec == 4 && from_logits != 0 && x->dtype == ECO_F32 && (int64_t)N * HW <= ((int64_t)1 << 31);
}
// third-generation fused step: logits (and the gradient) fp32 or bf16, labels fp32 or bytes; every plane and every tile of
// it must start on a 16-byte boundary and be a multiple of 16 bytes long (1-D TMA bulk copies)
static bool tma_planes_ok(const void* ptr, int64_t sn, int64_t sc, int dtype, int64_t HW) {
    const int64_t esz = dtype == ECO_F32 ? 4 : (dtype == ECO_BF16 ? 2 : 1), per16 = 16 / esz;
    return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && sn % per16 == 0 && sc % per16 == 0 && HW % per16 == 0;
}
// (probabilities -- ECO_C3_PROBS -- in fp32 only; `gx` may be null for a step without gradient)
static bool v3_eligible(const EcoView* x, const EcoView* g, const EcoOut* gx, bool from_logits, int32_t N, int64_t HW) {
    if ((int64_t)N * HW > ((int64_t)1 << 31)) return false;
    if (x->dtype != ECO_F32 && !(x->dtype == ECO_BF16 && from_logits)) return false;
    if (g->dtype != ECO_F32 && g->dtype != ECO_U8) return false;
    return tma_planes_ok(x->ptr, x->sn, x->sc, x->dtype, HW) && tma_planes_ok(g->ptr, g->sn, g->sc, g->dtype, HW) &&
           (!gx || tma_planes_ok(gx->ptr, gx->sn, gx->sc, gx->dtype, HW)) && HW % 4 == 0;
}
static int v2_grid(int device, int32_t N, int64_t HW) {
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -1;
    const int64_t tiles = (int64_t)N * ((HW + v2::kTP - 1) / v2::kTP);
    int64_t g = sms;   // one CTA per SM: co-resident, the grid-wide hand-over spins
    if (g > tiles) g = tiles;
    if (g > v2::kMaxGrid) g = v2::kMaxGrid;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace eco

using namespace eco;

// workspace: [0,256) arrival counters / status | 128 doubles of totals | per-CTA partials | v2 integer accumulators
static constexpr int64_t kWsV2Offset = 256 + 128 * 8 + (int64_t)kMaxCompCtas * kNAcc * (int64_t)sizeof(double);
static constexpr int64_t kWsV3Offset = (kWsV2Offset + (int64_t)sizeof(v2::V2Ws) + 255) / 256 * 256;
extern "C" int64_t eco_composite3_ws_bytes(void) { return kWsV3Offset + (int64_t)sizeof(v2::V3Ws); }

// scalar kernels serve the unaligned / ragged path (VEC == 1); the aligned path runs the packed kernels
#define ECO_DISPATCH_SCALAR(KERNEL, xdt, logits, ...)                                            \
    do {                                                                                          \
        if (xdt == ECO_F32) { if (logits) KERNEL<float, 1, true> __VA_ARGS__; else KERNEL<float, 1, false> __VA_ARGS__; } \
        else { if (logits) KERNEL<__nv_bfloat16, 1, true> __VA_ARGS__; else KERNEL<__nv_bfloat16, 1, false> __VA_ARGS__; } \
    } while (0)
#define ECO_DISPATCH_PACKED(KERNEL, xdt, logits, ...)                                            \
    do {                                                                                          \
        if (xdt == ECO_F32) { if (logits) KERNEL<float, true> __VA_ARGS__; else KERNEL<float, false> __VA_ARGS__; } \
        else { if (logits) KERNEL<__nv_bfloat16, true> __VA_ARGS__; else KERNEL<__nv_bfloat16, false> __VA_ARGS__; } \
    } while (0)

extern "C" int eco_composite3_stats(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                                    void* ws, int64_t ws_bytes, double* acc_out, int device, void* stream) {
    int rc = check_comp(x, g, N, HW);
    if (rc) return rc;
    if ((rc = need_f32_labels(g, "eco_composite3_stats"))) return rc;
    if (!ws || ws_bytes < eco_composite3_ws_bytes() || !acc_out) { set_error("workspace too small or null output"); return -5; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const int vec = (c_aligned(x->ptr, x->sn, x->sc, x->dtype, HW) && c_aligned(g->ptr, g->sn, g->sc, g->dtype, HW)) ? 4 : 1;
    CompArgs a{};
    fill_comp(a, x, g, N, HW, vec);
    const int grid = comp_grid(device, a.units_total, 1, vec == 4 ? kRoleThreads : kCThreads);
    if (grid < 0) return -10;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256 + 128 * 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec == 4) {
        rc = ensure_packed_smem();
        if (rc) return rc;
        ECO_DISPATCH_PACKED(composite3_stats_packed_kernel, x->dtype, from_logits != 0, <<<grid, kPThreads, kStageBytes, st>>>(a, counter, partials, acc_out));
    }
    else ECO_DISPATCH_SCALAR(composite3_stats_kernel, x->dtype, from_logits != 0, <<<grid, kCThreads, 0, st>>>(a, counter, partials, acc_out));
    return check_cuda(cudaGetLastError(), "composite3_stats kernel launch");
}

extern "C" int eco_composite3_finalize(const double* acc, const double* leaf_scale_host, const double* leaf_scale_dev,
                                       float* losses_out, double* jac_out, double* leaf_sums_out, int device,
                                       void* stream) {
    if (!acc || (!leaf_scale_host && !leaf_scale_dev)) { set_error("null acc / scales"); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    CompFinArgs fa{};
    if (leaf_scale_host)
        for (int l = 0; l < ECO_C3_NLEAF; ++l) fa.scale[l] = leaf_scale_host[l];
    composite3_finalize_kernel<<<1, 160, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        acc, fa, leaf_scale_host ? nullptr : leaf_scale_dev, losses_out, jac_out, leaf_sums_out);
    return check_cuda(cudaGetLastError(), "composite3_finalize_kernel launch");
}

extern "C" int eco_composite3_grad(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                                   const double* jac, const float* upstream, const EcoOut* gx, int device,
                                   void* stream) {
    int rc = check_comp(x, g, N, HW);
    if (rc) return rc;
    if ((rc = need_f32_labels(g, "eco_composite3_grad"))) return rc;
    if (!jac || !upstream || !gx || !gx->ptr) { set_error("null jac/upstream/gx"); return -5; }
    if (gx->dtype != x->dtype) { set_error("gx dtype must match x"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const int vec = (c_aligned(x->ptr, x->sn, x->sc, x->dtype, HW) && c_aligned(g->ptr, g->sn, g->sc, g->dtype, HW) &&
                     c_aligned(gx->ptr, gx->sn, gx->sc, gx->dtype, HW)) ? 4 : 1;
    CompGradArgs ga{};
    fill_comp(ga.a, x, g, N, HW, vec);
    ga.gx = gx->ptr; ga.gx_sn = gx->sn; ga.gx_sc = gx->sc;
    const int grid = comp_grid(device, ga.a.units_total, vec == 4 ? 1 : 2, vec == 4 ? kGThreads : kCThreads);
    if (grid < 0) return -10;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (v2_eligible(x, from_logits, vec, N, HW)) {
        rc = ensure_packed_smem();
        if (rc) return rc;
        const int g2 = v2_grid(device, N, HW);
        if (g2 < 0) return -10;
        v2::composite3_grad_v2_kernel<<<g2, v2::kThreads, v2::kSmemBytes, st>>>(ga, jac, upstream);
        return check_cuda(cudaGetLastError(), "composite3_grad_v2_kernel launch");
    }
    if (vec == 4) {
        rc = ensure_packed_smem();
        if (rc) return rc;
        ECO_DISPATCH_PACKED(composite3_grad_packed_kernel, x->dtype, from_logits != 0, <<<grid, kGThreads, kStageBytes, st>>>(ga, jac, upstream));
    }
    else ECO_DISPATCH_SCALAR(composite3_grad_kernel, x->dtype, from_logits != 0, <<<grid, kCThreads, 0, st>>>(ga, jac, upstream));
    return check_cuda(cudaGetLastError(), "composite3_grad kernel launch");
}

static int launch_fused(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, uint32_t flags,
                        const double* leaf_scale_dev, const float* upstream, void* ws, int64_t ws_bytes,
                        float* losses_out, const EcoOut* gx, XchArgs xch, int device, void* stream,
                        const float* upstream_prev = nullptr) {
    int rc = check_comp(x, g, N, HW);
    if (rc) return rc;
    if (flags & ~(uint32_t)(ECO_C3_UNION_LABELS | ECO_C3_PROBS | ECO_C3_NO_GRAD)) { set_error("unknown flags 0x%x", flags); return -3; }
    const bool no_grad = (flags & ECO_C3_NO_GRAD) != 0;
    if (no_grad && (!gx || !gx->ptr)) gx = nullptr;
    if (!leaf_scale_dev || !upstream || !losses_out || (!no_grad && (!gx || !gx->ptr))) { set_error("null scale/upstream/output"); return -5; }
    if (!ws || ws_bytes < eco_composite3_ws_bytes()) { set_error("workspace too small"); return -5; }
    if (gx && gx->dtype != x->dtype) { set_error("gx dtype must match x"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const bool lg = (flags & ECO_C3_PROBS) == 0;
    const bool g_f32 = g->dtype == ECO_F32;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    if (!xch.status) xch.status = counter + 32;
    if (v3_eligible(x, g, gx, lg, N, HW)) {
        CompGradArgs ga{};
        fill_comp(ga.a, x, g, N, HW, 4);
        if (gx) { ga.gx = gx->ptr; ga.gx_sn = gx->sn; ga.gx_sc = gx->sc; }
        rc = ensure_packed_smem();
        if (rc) return rc;
        const int g2 = v2_grid(device, N, HW);
        if (g2 < 0) return -10;
        v2::V3Ws* ws3 = reinterpret_cast<v2::V3Ws*>(reinterpret_cast<char*>(ws) + kWsV3Offset);
        void* args[] = {&ga, (void*)&leaf_scale_dev, (void*)&upstream, &ws3, &losses_out, &flags, &xch, (void*)&upstream_prev};
        const void* fn;
        int smem;
#define ECO_V3_PICK(TX, TG, PR) do { fn = (const void*)v2::composite3_fused_v3_kernel<TX, TG, PR>; smem = v2::Stage3<TX, TG>::kSmem; } while (0)
        if (!lg) { if (g_f32) ECO_V3_PICK(float, float, true); else ECO_V3_PICK(float, uint8_t, true); }
        else if (x->dtype == ECO_F32) { if (g_f32) ECO_V3_PICK(float, float, false); else ECO_V3_PICK(float, uint8_t, false); }
        else { if (g_f32) ECO_V3_PICK(__nv_bfloat16, float, false); else ECO_V3_PICK(__nv_bfloat16, uint8_t, false); }
#undef ECO_V3_PICK
        return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(g2), dim3(v2::kThreads3), args, smem, st), "composite3_fused_v3_kernel launch");
    }
    if (no_grad) { set_error("ECO_C3_NO_GRAD needs fp32 inputs (or bf16 logits) with 16-byte aligned planes; use eco_composite3_stats + eco_composite3_finalize otherwise"); return -8; }
    CompGradArgs ga{};
    ga.gx = gx->ptr; ga.gx_sn = gx->sn; ga.gx_sc = gx->sc;
    double* acc_glob = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    double* partials = acc_glob + 128;
    if (upstream_prev) { set_error("eco_composite3_step_if_changed needs 16-byte aligned planes"); return -8; }
    // everything below: the scalar first-generation kernel (bf16 probabilities, ragged / unaligned planes), f32 labels only
    if (!g_f32 || (flags & ECO_C3_UNION_LABELS)) {
        set_error("byte labels / the fused label union need logits with 16-byte aligned planes (H*W %% 16 == 0 for byte labels, %% 8 for bf16 logits, %% 4 otherwise)");
        return -8;
    }
    if (xch.world > 1) { set_error("the peer-exchange fused step needs fp32 / bf16 logits or fp32 probabilities with 16-byte aligned planes"); return -8; }
    // scalar kernel: any alignment, any supported dtype (one thread = one pixel)
    fill_comp(ga.a, x, g, N, HW, 1);
    const int grid = comp_grid(device, ga.a.units_total, 1, kCThreads);  // one CTA per SM: co-resident
    if (grid < 0) return -10;
    void* args[] = {&ga, (void*)&leaf_scale_dev, (void*)&upstream, &counter, &partials, &acc_glob, &losses_out};
    const void* fn;
    if (x->dtype == ECO_F32) fn = lg ? (const void*)composite3_fused_kernel<float, 1, true> : (const void*)composite3_fused_kernel<float, 1, false>;
    else fn = lg ? (const void*)composite3_fused_kernel<__nv_bfloat16, 1, true> : (const void*)composite3_fused_kernel<__nv_bfloat16, 1, false>;
    return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kCThreads), args, 0, st), "composite3_fused_kernel launch");
}

static constexpr double kDefaultXchTimeoutMs = 30000.0;

static int fill_xch(XchArgs& xch, void* const* peer_xch_dev, int32_t rank, int32_t world, uint32_t epoch, uint32_t* status,
                    double timeout_ms) {
    if (!peer_xch_dev || world < 1 || world > 64 || rank < 0 || rank >= world || epoch == 0) {
        set_error("bad exchange arguments (world=%d rank=%d epoch=%u)", world, rank, epoch);
        return -9;
    }
    xch.peers = reinterpret_cast<double* const*>(peer_xch_dev);
    xch.rank = rank;
    xch.world = world;
    xch.epoch = epoch;
    xch.status = status;
    if (!(timeout_ms > 0.0)) timeout_ms = kDefaultXchTimeoutMs;
    xch.timeout_ns = (unsigned long long)(timeout_ms * 1e6);
    return 0;
}

extern "C" int eco_composite3_step(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, uint32_t flags,
                                   const double* leaf_scale_dev, const float* upstream, void* ws, int64_t ws_bytes,
                                   float* losses_out, const EcoOut* gx, const EcoPeerExchange* peers, int device,
                                   void* stream) {
    XchArgs xch{};
    xch.world = 1;
    if (peers) {
        int rc = fill_xch(xch, peers->peer_xch_dev, peers->rank, peers->world, peers->epoch, peers->status, peers->timeout_ms);
        if (rc) return rc;
    }
    return launch_fused(x, g, N, HW, flags, leaf_scale_dev, upstream, ws, ws_bytes, losses_out, gx, xch, device, stream);
}

extern "C" int eco_composite3_step_if_changed(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, uint32_t flags,
                                              const double* leaf_scale_dev, const float* upstream, const float* upstream_prev,
                                              void* ws, int64_t ws_bytes, float* losses_out, const EcoOut* gx, int device,
                                              void* stream) {
    if (!upstream_prev) { set_error("null upstream_prev"); return -5; }
    XchArgs xch{};
    xch.world = 1;
    return launch_fused(x, g, N, HW, flags, leaf_scale_dev, upstream, ws, ws_bytes, losses_out, gx, xch, device, stream, upstream_prev);
}

extern "C" int eco_composite3_fused(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                                    const double* leaf_scale_dev, const float* upstream, void* ws, int64_t ws_bytes,
                                    float* losses_out, const EcoOut* gx, int device, void* stream) {
    XchArgs xch{};
    xch.world = 1;
    return launch_fused(x, g, N, HW, from_logits ? 0u : ECO_C3_PROBS, leaf_scale_dev, upstream, ws, ws_bytes, losses_out, gx, xch, device, stream);
}

extern "C" int eco_composite3_fused_sharded(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                                            const double* leaf_scale_dev, const float* upstream, void* ws,
                                            int64_t ws_bytes, float* losses_out, const EcoOut* gx,
                                            void* const* peer_xch_dev, int32_t rank, int32_t world, uint32_t epoch,
                                            int device, void* stream) {
    XchArgs xch{};
    int rc = fill_xch(xch, peer_xch_dev, rank, world, epoch, nullptr, 0.0);
    if (rc) return rc;
    return launch_fused(x, g, N, HW, from_logits ? 0u : ECO_C3_PROBS, leaf_scale_dev, upstream, ws, ws_bytes, losses_out, gx, xch, device, stream);
}

// Status word of the peer exchange inside `ws` (set by a kernel whose wait on a peer timed out): copies it to the host,
// clears it when set.  Synchronises `stream`.
extern "C" int eco_xch_poll_status(void* ws, int64_t ws_bytes, uint32_t* status_out_host, int device, void* stream) {
    if (!ws || ws_bytes < eco_composite3_ws_bytes() || !status_out_host) { set_error("bad eco_xch_poll_status arguments"); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* word = reinterpret_cast<unsigned int*>(ws) + 32;
    ECO_CUDA(cudaMemcpyAsync(status_out_host, word, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ECO_CUDA(cudaStreamSynchronize(st));
    if (*status_out_host) ECO_CUDA(cudaMemsetAsync(word, 0, sizeof(uint32_t), st));
    return 0;
}

extern "C" int eco_multiclass3_step(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, double leaf_scale,
                                    const float* upstream, const float* upstream_prev, uint32_t flags, void* ws,
                                    int64_t ws_bytes, float* losses_out, const EcoOut* gx, int device, void* stream) {
    int rc = check_comp(x, g, N, HW);
    if (rc) return rc;
    if ((rc = need_f32_labels(g, "eco_multiclass3_step"))) return rc;
    if (flags & ~(uint32_t)ECO_C3_PROBS) { set_error("eco_multiclass3_step: unknown flag bits 0x%x", flags); return -4; }
    if (!upstream || !losses_out || !gx || !gx->ptr) { set_error("null upstream/output"); return -5; }
    if (!ws || ws_bytes < eco_composite3_ws_bytes()) { set_error("workspace too small"); return -5; }
    if (gx->dtype != x->dtype) { set_error("gx dtype must match x"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const int vec = (c_aligned(x->ptr, x->sn, x->sc, x->dtype, HW) && c_aligned(g->ptr, g->sn, g->sc, g->dtype, HW) &&
                     c_aligned(gx->ptr, gx->sn, gx->sc, gx->dtype, HW)) ? 4 : 1;
    if (!v2_eligible(x, 1, vec, N, HW)) {
        set_error("eco_multiclass3_step serves fp32 logits / probabilities with 16-byte aligned planes (H*W %% 4 == 0); use eco_pair_* otherwise");
        return -8;
    }
    CompGradArgs ga{};
    fill_comp(ga.a, x, g, N, HW, vec);
    ga.gx = gx->ptr; ga.gx_sn = gx->sn; ga.gx_sc = gx->sc;
    rc = ensure_packed_smem();
    if (rc) return rc;
    const int grid = v2_grid(device, N, HW);
    if (grid < 0) return -10;
    v2::V2Ws* ws2 = reinterpret_cast<v2::V2Ws*>(reinterpret_cast<char*>(ws) + kWsV2Offset);
    void* args[] = {&ga, &leaf_scale, (void*)&upstream, &ws2, &losses_out, (void*)&upstream_prev};
    const void* kernel = (flags & ECO_C3_PROBS) ? (const void*)v2::multiclass3_fused_v2_kernel<true>
                                                : (const void*)v2::multiclass3_fused_v2_kernel<false>;
    return check_cuda(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(v2::kThreads), args, v2::kSmemBytes,
                                                  reinterpret_cast<cudaStream_t>(stream)),
                      "multiclass3_fused_v2_kernel launch");
}

extern "C" int eco_multiclass3_fused(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, double leaf_scale,
                                     const float* upstream, void* ws, int64_t ws_bytes, float* losses_out,
                                     const EcoOut* gx, int device, void* stream) {
    return eco_multiclass3_step(x, g, N, HW, leaf_scale, upstream, nullptr, 0u, ws, ws_bytes, losses_out, gx, device, stream);
}

// ---- peer exchange buffers (CUDA IPC).  The one place the library allocates: IPC needs a cudaMalloc base pointer. ----
extern "C" int64_t eco_xch_bytes(int32_t world) {
    if (world < 1 || world > 64) return -1;
    return (int64_t)v2::xch_ll_offset_bytes(world) + (int64_t)2 * world * 128 * 16;  // first-generation slots + flags, then the LL rows
}

extern "C" int eco_xch_alloc(int32_t world, void** ptr_out, unsigned char* handle_out /*[64]*/, int device) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ptr_out || !handle_out || eco_xch_bytes(world) < 0) { set_error("bad eco_xch_alloc arguments"); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    void* p = nullptr;
    ECO_CUDA(cudaMalloc(&p, (size_t)eco_xch_bytes(world)));
    ECO_CUDA(cudaMemset(p, 0, (size_t)eco_xch_bytes(world)));
    ECO_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    int rc = check_cuda(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    if (rc) { cudaFree(p); return rc; }
    memcpy(handle_out, &h, 64);
    *ptr_out = p;
    return 0;
}

extern "C" int eco_xch_open(const unsigned char* handle /*[64]*/, void** ptr_out, int device) {
    if (!handle || !ptr_out) { set_error("bad eco_xch_open arguments"); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return check_cuda(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

extern "C" int eco_xch_close(void* peer_ptr, int device) {
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    return check_cuda(cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle");
}

extern "C" int eco_xch_free(void* ptr, int device) {
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    return check_cuda(cudaFree(ptr), "cudaFree");
}
