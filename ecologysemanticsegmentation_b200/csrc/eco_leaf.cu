// Pair-leaf engine: C independent (a_c, b_c) leaves over N images, one streaming pass for the
// statistics and one for the gradient.  See include/ecoloss.h for the contract.
//
// HBM-bound design (B200, 148 SMs): 256-thread CTAs, each thread keeps UNROLL independent 128-bit
// loads per operand in flight, a contiguous run of tiles per CTA sized so the whole grid is one
// resident wave, fp32 per-tile partials folded into fp64 thread accumulators, warp-shuffle +
// shared-memory tree per CTA, deterministic last-CTA-per-channel reduction of the CTA partials.
#include <stdarg.h>
#include <stdio.h>

#include "eco_common.cuh"

namespace eco {

#ifdef ECO_LEAF_TIMELINE   // exp/leaf_timeline.py: %globaltimer stamps per CTA (not compiled into the library)
__device__ unsigned long long g_leaf_tl[4096 * 8];
__device__ __forceinline__ unsigned long long leaf_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ECO_LTL(slot) do { if (threadIdx.x == 0) g_leaf_tl[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = leaf_gtime(); } while (0)
#else
#define ECO_LTL(slot) do { } while (0)
#endif

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kCtasPerSm = 4;

struct PairArgs {
    const void* a;
    const void* b;
    int64_t a_sn, a_sc, b_sn, b_sc;
    int32_t N, C;
    int64_t HW;
    uint32_t flags;
    int32_t tiles_per_plane;
    int64_t tiles_per_channel;
    int32_t tiles_per_cta;
    float focal_gamma;  // read by the GEN instantiations only
};

// ws layout: [0, 64*C) bytes: one uint32 arrival counter per channel (padded);
//            then partials: double[C][max_ctas_per_channel][8]
constexpr int kMaxCtasPerChannel = 148 * kCtasPerSm * 2;

__host__ __device__ inline int64_t pair_ws_partials_offset(int C) { return ((int64_t)C * 4 + 255) / 256 * 256; }

// GEN: caller-chosen focal exponent (powf path); the default instantiation keeps the sqrt closed form.
// Pass 1 of one CTA: streams its tiles, reduces over the CTA, parks the 7 partial sums, arrives; the last CTA of the
// channel adds the partials of all CTAs in a fixed order and writes the channel's sums.  Returns true in that CTA
// (uniformly over its threads).  `rearm`: reset the arrival counter for the next launch on this workspace.
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

template <typename TA, typename TB, int VEC, bool GEN>
__device__ __forceinline__ bool pair_stats_cta(const PairArgs& p, unsigned int* __restrict__ counters,
                                               double* __restrict__ partials, double* __restrict__ sums_out,
                                               double* sums_sm = nullptr /* shared-memory copy, [ECO_NSTAT] */) {
    constexpr int kTile = kThreads * VEC * kUnroll;
    const int c = blockIdx.y;
    const bool a_logit = p.flags & ECO_A_LOGIT, b_logit = p.flags & ECO_B_LOGIT;
    const bool need_bg = p.flags & ECO_NEED_BG;
    const TA* __restrict__ abase = reinterpret_cast<const TA*>(p.a) + (int64_t)c * p.a_sc;
    const TB* __restrict__ bbase = reinterpret_cast<const TB*>(p.b) + (int64_t)c * p.b_sc;

    double dacc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) dacc[k] = 0.0;

    int64_t tile = (int64_t)blockIdx.x * p.tiles_per_cta;
    int64_t tile_end = tile + p.tiles_per_cta;
    if (tile_end > p.tiles_per_channel) tile_end = p.tiles_per_channel;
    int64_t n = tile / p.tiles_per_plane;
    int32_t t = (int32_t)(tile - n * p.tiles_per_plane);

    for (; tile < tile_end; ++tile) {
        const TA* ap = abase + n * p.a_sn;
        const TB* bp = bbase + n * p.b_sn;
        const int64_t e0 = (int64_t)t * kTile + (int64_t)threadIdx.x * VEC;
        float av[kUnroll][VEC], bv[kUnroll][VEC];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t e = e0 + (int64_t)u * kThreads * VEC;
            ok[u] = e < p.HW;
            if (ok[u]) {
                if constexpr (VEC == 4) {
                    Vec4<TA>::load(ap + e, reinterpret_cast<float(&)[4]>(av[u]));
                    Vec4<TB>::load(bp + e, reinterpret_cast<float(&)[4]>(bv[u]));
                } else {
                    av[u][0] = Vec4<TA>::load1(ap + e);
                    bv[u][0] = Vec4<TB>::load1(bp + e);
                }
            }
        }
        float acc[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] = 0.f;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (ok[u]) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float a = av[u][v], b = bv[u][v];
                    // no |.| kink or threshold on this path: the 2-ulp MUFU sigmoid is ample for sums at 1e-5
                    if (a_logit) a = sigmoid_fast(a);
                    if (b_logit) b = sigmoid_fast(b);
                    acc[0] += a;
                    acc[1] += b;
                    acc[2] = fmaf(a, b, acc[2]);
                    acc[3] = fmaf(b, b, acc[3]);
                    acc[4] += softplus_slot(b);
                    if constexpr (GEN) {
                        acc[5] -= focal_fg_log2_gen(b, p.focal_gamma);
                        if (need_bg) acc[6] -= focal_bg_log2_gen(b, p.focal_gamma);
                    } else {
                        acc[5] -= focal_fg_log2(b);
                        if (need_bg) acc[6] -= focal_bg_log2(b);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) dacc[k] += (double)acc[k];
        if (++t == p.tiles_per_plane) {
            t = 0;
            ++n;
        }
    }
    ECO_LTL(1);
    // focal terms were accumulated in log2 units
    dacc[5] *= kLn2d;
    dacc[6] *= kLn2d;

    // CTA reduction
    __shared__ double sm[kThreads / 32][7];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double v = warp_sum(dacc[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    double* my_partials = partials + ((int64_t)c * kMaxCtasPerChannel + blockIdx.x) * 8;
    if (threadIdx.x < 7) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += sm[w][threadIdx.x];
        my_partials[threadIdx.x] = v;
    }
    // Arrival without stand-alone fences (a sequentially-consistent MEMBAR costs ~1 us here, twice): the partial
    // stores of threads 0..6 precede thread 0's RELEASE through the CTA barrier (release is cumulative), and the last
    // CTA's reads below follow thread 0's ACQUIRE through the next barrier; they go to L2 (__ldcg), not to a stale L1.
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atom_add_acq_rel_gpu(&counters[c], 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    // deterministic final reduction by the last CTA of this channel: warp k sums stat k
    if (warp < 7) {
        const double* base = partials + (int64_t)c * kMaxCtasPerChannel * 8;
        double v = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(base + (int64_t)i * 8 + warp);
        v = warp_sum(v);
        if (lane == 0) {
            sums_out[c * ECO_NSTAT + 1 + warp] = v;
            if (sums_sm) sums_sm[1 + warp] = v;
        }
    }
    if (threadIdx.x == 0) {
        sums_out[c * ECO_NSTAT + S_N] = (double)p.N * (double)p.HW;
        if (sums_sm) sums_sm[S_N] = (double)p.N * (double)p.HW;
        counters[c] = 0;  // re-arm for the next launch on this workspace
    }
    return true;
}

template <typename TA, typename TB, int VEC, bool GEN>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
pair_stats_kernel(PairArgs p, unsigned int* __restrict__ counters, double* __restrict__ partials,
                  double* __restrict__ sums_out) {
    pair_stats_cta<TA, TB, VEC, GEN>(p, counters, partials, sums_out);
}

// ---------------------------------------------------------------------------------------------
// finalize: sums -> losses, Jacobian
// ---------------------------------------------------------------------------------------------
struct FinalizeArgs {
    double bw;
    double scale[64];
    LeafShape shape;
    int32_t shaped;
};

__global__ void pair_finalize_kernel(const double* __restrict__ sums, int C, FinalizeArgs fa,
                                     float* __restrict__ losses_out, float* __restrict__ total_out,
                                     double* __restrict__ jac_out) {
    __shared__ double sl[64][ECO_NLOSS];
    const int c = threadIdx.x;
    if (c < C) {
        LeafOut o;
        if (fa.shaped) leaf_closed_form(sums + c * ECO_NSTAT, fa.bw, fa.scale[c], fa.shape, o);
        else leaf_closed_form(sums + c * ECO_NSTAT, fa.bw, fa.scale[c], o);
        for (int k = 0; k < ECO_NLOSS; ++k) {
            sl[c][k] = o.loss[k];
            if (losses_out) losses_out[c * ECO_NLOSS + k] = (float)o.loss[k];
            if (jac_out)
                for (int j = 0; j < ECO_NJAC; ++j) jac_out[(c * ECO_NLOSS + k) * ECO_NJAC + j] = o.jac[k][j];
        }
    }
    __syncthreads();
    if (total_out && threadIdx.x < ECO_NLOSS) {
        double v = 0.0;
        for (int i = 0; i < C; ++i) v += sl[i][threadIdx.x];
        total_out[threadIdx.x] = (float)v;
    }
}

// ---------------------------------------------------------------------------------------------
// gradient pass
// ---------------------------------------------------------------------------------------------
struct GradArgs {
    PairArgs p;
    void* ga;
    void* gb;
    int64_t ga_sn, ga_sc, gb_sn, gb_sc;
    int32_t accumulate;
};

// Pass 2 of one CTA over its tiles with the channel's coefficients.
template <typename TA, typename TB, int VEC, bool GEN>
__device__ __forceinline__ void pair_grad_cta(const GradArgs& g, const LeafCoef cf) {
    constexpr int kTile = kThreads * VEC * kUnroll;
    const PairArgs& p = g.p;
    const int c = blockIdx.y;
    const bool a_logit = p.flags & ECO_A_LOGIT, b_logit = p.flags & ECO_B_LOGIT;
    const TA* __restrict__ abase = reinterpret_cast<const TA*>(p.a) + (int64_t)c * p.a_sc;
    const TB* __restrict__ bbase = reinterpret_cast<const TB*>(p.b) + (int64_t)c * p.b_sc;
    TA* gabase = g.ga ? reinterpret_cast<TA*>(g.ga) + (int64_t)c * g.ga_sc : nullptr;
    TB* gbbase = g.gb ? reinterpret_cast<TB*>(g.gb) + (int64_t)c * g.gb_sc : nullptr;
    const bool need_sig = cf.sp != 0.f, need_fl = cf.fl != 0.f, need_flb = cf.flb != 0.f;

    int64_t tile = (int64_t)blockIdx.x * p.tiles_per_cta;
    int64_t tile_end = tile + p.tiles_per_cta;
    if (tile_end > p.tiles_per_channel) tile_end = p.tiles_per_channel;
    int64_t n = tile / p.tiles_per_plane;
    int32_t t = (int32_t)(tile - n * p.tiles_per_plane);

    for (; tile < tile_end; ++tile) {
        const TA* ap = abase + n * p.a_sn;
        const TB* bp = bbase + n * p.b_sn;
        TA* gap = gabase ? gabase + n * g.ga_sn : nullptr;
        TB* gbp = gbbase ? gbbase + n * g.gb_sn : nullptr;
        const int64_t e0 = (int64_t)t * kTile + (int64_t)threadIdx.x * VEC;
        float av[kUnroll][VEC], bv[kUnroll][VEC];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t e = e0 + (int64_t)u * kThreads * VEC;
            ok[u] = e < p.HW;
            if (ok[u]) {
                if constexpr (VEC == 4) {
                    Vec4<TA>::load(ap + e, reinterpret_cast<float(&)[4]>(av[u]));
                    Vec4<TB>::load(bp + e, reinterpret_cast<float(&)[4]>(bv[u]));
                } else {
                    av[u][0] = Vec4<TA>::load1(ap + e);
                    bv[u][0] = Vec4<TB>::load1(bp + e);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!ok[u]) continue;
            const int64_t e = e0 + (int64_t)u * kThreads * VEC;
            float oa[VEC], ob[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float a = av[u][v], b = bv[u][v];
                if (a_logit) a = sigmoid_fast(a);
                if (b_logit) b = sigmoid_fast(b);
                float da = fmaf(cf.sab, b, cf.sa);
                float db = fmaf(cf.sab, a, fmaf(cf.sbb2, b, cf.sb));
                if (need_sig) db = fmaf(cf.sp, sigmoid_fast(b), db);
                if constexpr (GEN) {
                    if (need_fl) db = fmaf(cf.fl, dfocal_fg_gen(b, p.focal_gamma), db);
                    if (need_flb) db = fmaf(cf.flb, dfocal_bg_gen(b, p.focal_gamma), db);
                } else {
                    if (need_fl) db = fmaf(cf.fl, dfocal_fg(b), db);
                    if (need_flb) db = fmaf(cf.flb, dfocal_bg(b), db);
                }
                if (a_logit) da *= (1.0f - a) * a;
                if (b_logit) db *= (1.0f - b) * b;
                oa[v] = da;
                ob[v] = db;
            }
            if (gap) {
                if constexpr (VEC == 4) {
                    if (g.accumulate) {
                        float old[4];
                        Vec4<TA>::load(gap + e, old);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) oa[v] += old[v];
                    }
                    Vec4<TA>::store(gap + e, reinterpret_cast<float(&)[4]>(oa));
                } else {
                    if (g.accumulate) oa[0] += Vec4<TA>::load1(gap + e);
                    Vec4<TA>::store1(gap + e, oa[0]);
                }
            }
            if (gbp) {
                if constexpr (VEC == 4) {
                    if (g.accumulate) {
                        float old[4];
                        Vec4<TB>::load(gbp + e, old);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) ob[v] += old[v];
                    }
                    Vec4<TB>::store(gbp + e, reinterpret_cast<float(&)[4]>(ob));
                } else {
                    if (g.accumulate) ob[0] += Vec4<TB>::load1(gbp + e);
                    Vec4<TB>::store1(gbp + e, ob[0]);
                }
            }
        }
        if (++t == p.tiles_per_plane) {
            t = 0;
            ++n;
        }
    }
}

template <typename TA, typename TB, int VEC, bool GEN>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
pair_grad_kernel(GradArgs g, const double* __restrict__ jac, const float* __restrict__ upstream) {
    __shared__ LeafCoef coef_s;
    if (threadIdx.x == 0) coef_s = make_coef(jac + (int64_t)blockIdx.y * ECO_NLOSS * ECO_NJAC, upstream);
    __syncthreads();
    pair_grad_cta<TA, TB, VEC, GEN>(g, coef_s);
}

// ---------------------------------------------------------------------------------------------
// One cooperative launch for a whole step of C leaves: pass 1 -> per-channel hand-over (the last CTA of a channel
// forms the sums, the closed forms and the gradient coefficients and releases the channel's other CTAs) -> pass 2
// over the same tiles (mostly L2 hits).  All CTAs are co-resident (grid = one resident wave, cooperative launch).
// Workspace words after the stats layout: per channel {done generation u32}, the coefficients, the losses; one
// grid-wide counter of finished channels so that the last one adds up the 7 totals in channel order.
// ---------------------------------------------------------------------------------------------
struct FusedWs {
    unsigned int done[64];       // generation of the last completed hand-over per channel
    unsigned int chan_count;     // channels finalized in this launch
    unsigned int _pad[63];
    LeafCoef coef[64];
    double loss[64][ECO_NLOSS];
};

struct FusedArgs {
    GradArgs g;
    double bw, scale;
    LeafShape shape;
    int32_t shaped;
    const float* upstream_prev;   // "only if changed": the outputs already hold the step for these weights (or null)
};
// the backward half of the drop-in autograd path: every thread of every CTA compares the two weight vectors (same memory,
// same answer) and the whole grid leaves before it touches anything when they agree
__device__ __forceinline__ bool upstream_unchanged(const float* upstream, const float* upstream_prev) {
    if (!upstream_prev) return false;
    bool same = true;
#pragma unroll
    for (int k = 0; k < ECO_NLOSS; ++k) same = same && (__float_as_uint(upstream[k]) == __float_as_uint(upstream_prev[k]));
    return same;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// hand-over by the channel's last CTA (all its threads call this): warp k forms loss k and its Jacobian row -- one
// warp per loss kind, because the seven float64 closed forms are seven different code paths and would serialise inside
// one warp -- then seven threads contract the rows with the upstream weights into the gradient coefficients.  Kept
// out of line: the float64 code must not weigh on the register allocation of the two streaming loops.
__device__ __noinline__ void fused_handover(const FusedArgs& fa, const float* upstream /* shared [7] */, FusedWs* __restrict__ fw,
                                            const double* s /* shared: the channel's sums */, float* __restrict__ losses_out,
                                            int c, unsigned int gen) {
    __shared__ double sl[ECO_NLOSS];
    __shared__ double sj[ECO_NLOSS][ECO_NJAC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < ECO_NLOSS; k += kThreads / 32) {
        if (lane == 0) {
            if (fa.shaped) leaf_closed_form_row<false>(s, fa.bw, fa.scale, k, sl[k], sj[k], fa.shape);
            else leaf_closed_form_row(s, fa.bw, fa.scale, k, sl[k], sj[k]);
        }
    }
    __syncthreads();
    ECO_LTL(7);
    if (threadIdx.x < ECO_NJAC) {
        const int j = threadIdx.x;
        double v = 0.0;
#pragma unroll
        for (int k = 1; k < ECO_NLOSS; ++k) {
            const double w = (double)upstream[k];
            if (w != 0.0) v += w * sj[k][j];   // rows without upstream weight stay out (their Jacobian may be non-finite)
        }
        reinterpret_cast<float*>(&fw->coef[c])[j] = (float)(j == 3 ? 2.0 * v : v);   // LeafCoef order; [3] = 2 c_Sbb
        fw->loss[c][j] = sl[j];
        if (gridDim.y == 1) losses_out[j] = (float)sl[j];   // a single leaf: its losses are the totals
    }
    __syncthreads();   // those stores precede thread 0's release below (cumulative through the barrier)
    if (threadIdx.x == 0) st_release_u32(&fw->done[c], gen + 1u);
}

// Second half of the hand-over, off the critical path: ONE thread of the channel's last CTA, after its pass 2.  The
// last channel to get here adds the 7 totals in channel order (deterministic).
__device__ __forceinline__ void fused_totals(FusedWs* __restrict__ fw, float* __restrict__ losses_out) {
    const unsigned int prev = atom_add_acq_rel_gpu(&fw->chan_count, 1u);
    if (prev == gridDim.y - 1) {
        for (int k = 0; k < ECO_NLOSS; ++k) {
            double v = 0.0;
            for (int cc = 0; cc < (int)gridDim.y; ++cc) v += __ldcg(&fw->loss[cc][k]);
            losses_out[k] = (float)v;
        }
        fw->chan_count = 0u;
    }
}

template <typename TA, typename TB, int VEC, bool GEN>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
pair_fused_kernel(FusedArgs fa, const float* __restrict__ upstream, unsigned int* __restrict__ counters,
                  double* __restrict__ partials, FusedWs* __restrict__ fw, double* __restrict__ sums_out,
                  float* __restrict__ losses_out) {
    if (upstream_unchanged(upstream, fa.upstream_prev)) return;
    const int c = blockIdx.y;
    __shared__ unsigned int gen_s;
    __shared__ LeafCoef coef_s;
    __shared__ float up_s[ECO_NLOSS + 1];
    // fetched now, under pass 1: nothing on the hand-over path between the two passes waits for a global load
    if (threadIdx.x >= 32 && threadIdx.x < 32 + ECO_NLOSS) up_s[threadIdx.x - 32] = upstream[threadIdx.x - 32];
    // the generation is read BEFORE this CTA arrives: the hand-over of this launch cannot have happened yet
    if (threadIdx.x == 0) gen_s = ld_acquire_u32(&fw->done[c]);
    __syncthreads();
    const unsigned int gen = gen_s;
    ECO_LTL(0);
    __shared__ double sums_s[ECO_NSTAT];
    const bool last = pair_stats_cta<TA, TB, VEC, GEN>(fa.g.p, counters, partials, sums_out, sums_s);
    ECO_LTL(2);
    if (last) {
        __syncthreads();   // the channel's sums are in sums_s (and in sums_out for the caller)
        ECO_LTL(5);
        fused_handover(fa, up_s, fw, sums_s, losses_out, c, gen);
        ECO_LTL(6);
    }
    if (threadIdx.x == 0) {
        while (ld_acquire_u32(&fw->done[c]) == gen) __nanosleep(20);
        for (int j = 0; j < 7; ++j)
            reinterpret_cast<float*>(&coef_s)[j] = __ldcg(reinterpret_cast<const float*>(&fw->coef[c]) + j);
    }
    __syncthreads();
    ECO_LTL(3);
    pair_grad_cta<TA, TB, VEC, GEN>(fa.g, coef_s);
    if (last && threadIdx.x == 0 && gridDim.y > 1) fused_totals(fw, losses_out);
    ECO_LTL(4);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool aligned_for_vec4(const void* ptr, int64_t sn, int64_t sc, int dtype, int64_t HW) {
    const int64_t esz = dtype == ECO_BF16 ? 2 : 4;
    const int64_t need = 4 * esz;  // bytes per 4-vector
    return (reinterpret_cast<uintptr_t>(ptr) % need == 0) && (sn % 4 == 0) && (sc % 4 == 0) && (HW % 4 == 0);
}

static int plan(PairArgs& p, int vec, int device, dim3& grid) {
    const int tile = kThreads * vec * kUnroll;
    p.tiles_per_plane = (int32_t)((p.HW + tile - 1) / tile);
    p.tiles_per_channel = (int64_t)p.tiles_per_plane * p.N;
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -10;
    int64_t max_ctas = (int64_t)sms * kCtasPerSm / p.C;
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > kMaxCtasPerChannel) max_ctas = kMaxCtasPerChannel;
    int64_t per = (p.tiles_per_channel + max_ctas - 1) / max_ctas;
    if (per < 1) per = 1;
    p.tiles_per_cta = (int32_t)per;
    int64_t ctas = (p.tiles_per_channel + per - 1) / per;
    if (ctas < 1) ctas = 1;
    grid = dim3((unsigned)ctas, (unsigned)p.C, 1);
    return 0;
}

static int fill_args(PairArgs& p, const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW,
                     uint32_t flags) {
    if (N <= 0 || C <= 0 || HW <= 0) { set_error("empty input (N=%d C=%d HW=%lld)", N, C, (long long)HW); return -2; }
    if (!a || !b || !a->ptr || !b->ptr) { set_error("null input view"); return -1; }
    if (C > 65535) { set_error("C=%d exceeds grid.y limit", C); return -3; }
    if ((a->dtype != ECO_F32 && a->dtype != ECO_BF16) || (b->dtype != ECO_F32 && b->dtype != ECO_BF16)) {
        set_error("unsupported dtype code");
        return -4;
    }
    p.a = a->ptr; p.b = b->ptr;
    p.a_sn = a->sn; p.a_sc = a->sc; p.b_sn = b->sn; p.b_sc = b->sc;
    p.N = N; p.C = C; p.HW = HW; p.flags = flags;
    return 0;
}

#define ECO_DISPATCH_PAIR_G(KERNEL, adt, bdt, vec, GEN, ...)                                             \
    do {                                                                                                 \
        if (vec == 4) {                                                                                  \
            if (adt == ECO_F32 && bdt == ECO_F32) KERNEL<float, float, 4, GEN> __VA_ARGS__;              \
            else if (adt == ECO_BF16 && bdt == ECO_F32) KERNEL<__nv_bfloat16, float, 4, GEN> __VA_ARGS__; \
            else if (adt == ECO_F32 && bdt == ECO_BF16) KERNEL<float, __nv_bfloat16, 4, GEN> __VA_ARGS__; \
            else KERNEL<__nv_bfloat16, __nv_bfloat16, 4, GEN> __VA_ARGS__;                                \
        } else {                                                                                         \
            if (adt == ECO_F32 && bdt == ECO_F32) KERNEL<float, float, 1, GEN> __VA_ARGS__;              \
            else if (adt == ECO_BF16 && bdt == ECO_F32) KERNEL<__nv_bfloat16, float, 1, GEN> __VA_ARGS__; \
            else if (adt == ECO_F32 && bdt == ECO_BF16) KERNEL<float, __nv_bfloat16, 1, GEN> __VA_ARGS__; \
            else KERNEL<__nv_bfloat16, __nv_bfloat16, 1, GEN> __VA_ARGS__;                                \
        }                                                                                                \
    } while (0)
#define ECO_DISPATCH_PAIR(KERNEL, adt, bdt, vec, gen, ...)                       \
    do {                                                                         \
        if (gen) ECO_DISPATCH_PAIR_G(KERNEL, adt, bdt, vec, true, __VA_ARGS__);  \
        else ECO_DISPATCH_PAIR_G(KERNEL, adt, bdt, vec, false, __VA_ARGS__);     \
    } while (0)

// shape == NULL or the reference's default exponent -> the sqrt closed form; anything else -> powf instantiation
static bool general_focal(const EcoLeafShape* shape, float& gamma_out) {
    gamma_out = 1.5f;
    if (!shape) return false;
    gamma_out = (float)shape->focal_gamma;
    return shape->focal_gamma != 1.5;
}

}  // namespace eco

#include "eco_leaf_resident.cuh"

using namespace eco;

extern "C" int64_t eco_pair_ws_bytes(int32_t C) {
    if (C <= 0) return -1;
    return pair_ws_partials_offset(C) + (int64_t)C * kMaxCtasPerChannel * 8 * (int64_t)sizeof(double);
}

extern "C" int eco_pair_stats(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                              void* ws, int64_t ws_bytes, double* sums_out, int device, void* stream) {
    return eco_pair_stats_shaped(a, b, N, C, HW, flags, nullptr, ws, ws_bytes, sums_out, device, stream);
}

extern "C" int eco_pair_stats_shaped(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW,
                                     uint32_t flags, const EcoLeafShape* shape_host, void* ws, int64_t ws_bytes,
                                     double* sums_out, int device, void* stream) {
    PairArgs p{};
    const bool gen = general_focal(shape_host, p.focal_gamma);
    const float gamma = p.focal_gamma;
    int rc = fill_args(p, a, b, N, C, HW, flags);
    if (rc) return rc;
    if (!ws || ws_bytes < eco_pair_ws_bytes(C) || !sums_out) { set_error("workspace too small or null output"); return -5; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const int vec = (aligned_for_vec4(a->ptr, a->sn, a->sc, a->dtype, HW) &&
                     aligned_for_vec4(b->ptr, b->sn, b->sc, b->dtype, HW)) ? 4 : 1;
    dim3 grid;
    rc = plan(p, vec, device, grid);
    if (rc) return rc;
    p.focal_gamma = gamma;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + pair_ws_partials_offset(C));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ECO_DISPATCH_PAIR(pair_stats_kernel, a->dtype, b->dtype, vec, gen, <<<grid, kThreads, 0, st>>>(p, counters, partials, sums_out));
    return check_cuda(cudaGetLastError(), "pair_stats_kernel launch");
}

extern "C" int eco_pair_finalize(const double* sums, int32_t C, double background_weight, const double* scale_host,
                                 float* losses_out, float* total_out, double* jac_out, int device, void* stream) {
    return eco_pair_finalize_shaped(sums, C, background_weight, scale_host, nullptr, losses_out, total_out, jac_out,
                                    device, stream);
}

extern "C" int eco_pair_finalize_shaped(const double* sums, int32_t C, double background_weight,
                                        const double* scale_host, const EcoLeafShape* shape_host, float* losses_out,
                                        float* total_out, double* jac_out, int device, void* stream) {
    if (!sums || C <= 0 || C > 64) { set_error("eco_pair_finalize: C must be in [1,64] (got %d)", C); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    FinalizeArgs fa{};
    fa.bw = background_weight;
    for (int c = 0; c < C; ++c) fa.scale[c] = scale_host ? scale_host[c] : 1.0;
    if (shape_host && (shape_host->tversky_alpha != 0.5 || shape_host->tversky_beta != 0.3 ||
                       shape_host->focal_dice_gamma != 1.8)) {
        fa.shaped = 1;
        fa.shape.alpha = shape_host->tversky_alpha;
        fa.shape.beta = shape_host->tversky_beta;
        fa.shape.fd_gamma = shape_host->focal_dice_gamma;
    }
    pair_finalize_kernel<<<1, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sums, C, fa, losses_out, total_out, jac_out);
    return check_cuda(cudaGetLastError(), "pair_finalize_kernel launch");
}

extern "C" int eco_pair_grad(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                             const double* jac, const float* upstream, const EcoOut* ga, const EcoOut* gb,
                             int32_t accumulate, int device, void* stream) {
    return eco_pair_grad_shaped(a, b, N, C, HW, flags, nullptr, jac, upstream, ga, gb, accumulate, device, stream);
}

extern "C" int eco_pair_grad_shaped(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW,
                                    uint32_t flags, const EcoLeafShape* shape_host, const double* jac,
                                    const float* upstream, const EcoOut* ga, const EcoOut* gb, int32_t accumulate,
                                    int device, void* stream) {
    GradArgs g{};
    float gamma;
    const bool gen = general_focal(shape_host, gamma);
    int rc = fill_args(g.p, a, b, N, C, HW, flags);
    if (rc) return rc;
    if (!jac || !upstream) { set_error("null jac/upstream"); return -5; }
    const bool want_a = ga && ga->ptr, want_b = gb && gb->ptr;
    if (!want_a && !want_b) return 0;
    if (want_a && ga->dtype != a->dtype) { set_error("ga dtype must match slot a"); return -7; }
    if (want_b && gb->dtype != b->dtype) { set_error("gb dtype must match slot b"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    bool v4 = aligned_for_vec4(a->ptr, a->sn, a->sc, a->dtype, HW) && aligned_for_vec4(b->ptr, b->sn, b->sc, b->dtype, HW);
    if (want_a) { g.ga = ga->ptr; g.ga_sn = ga->sn; g.ga_sc = ga->sc; v4 = v4 && aligned_for_vec4(ga->ptr, ga->sn, ga->sc, ga->dtype, HW); }
    if (want_b) { g.gb = gb->ptr; g.gb_sn = gb->sn; g.gb_sc = gb->sc; v4 = v4 && aligned_for_vec4(gb->ptr, gb->sn, gb->sc, gb->dtype, HW); }
    g.accumulate = accumulate;
    const int vec = v4 ? 4 : 1;
    dim3 grid;
    rc = plan(g.p, vec, device, grid);
    if (rc) return rc;
    g.p.focal_gamma = gamma;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ECO_DISPATCH_PAIR(pair_grad_kernel, a->dtype, b->dtype, vec, gen, <<<grid, kThreads, 0, st>>>(g, jac, upstream));
    return check_cuda(cudaGetLastError(), "pair_grad_kernel launch");
}

// ------------------------------------------------------------------------------------------------------------------
// fused step
// ------------------------------------------------------------------------------------------------------------------
static int64_t fused_ws_offset(int32_t C) { return (eco_pair_ws_bytes(C) + 255) / 256 * 256; }

static int64_t resident_ws_offset(int32_t C) { return (fused_ws_offset(C) + (int64_t)sizeof(FusedWs) + 255) / 256 * 256; }

extern "C" int64_t eco_pair_fused_ws_bytes(int32_t C) {
    if (C <= 0 || C > 64) return -1;
    return resident_ws_offset(C) + (int64_t)sizeof(resident::ResidentWs);
}

extern "C" int eco_pair_fused_ex(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                                 double background_weight, double scale, const EcoLeafShape* shape_host,
                                 const float* upstream, const float* upstream_prev, void* ws, int64_t ws_bytes,
                                 double* sums_out, float* losses_out, const EcoOut* ga, const EcoOut* gb, int device,
                                 void* stream) {
    FusedArgs fa{};
    fa.upstream_prev = upstream_prev;
    GradArgs& g = fa.g;
    int rc = fill_args(g.p, a, b, N, C, HW, flags | (background_weight != 0.0 ? ECO_NEED_BG : 0u));
    if (rc) return rc;
    if (C > 64) { set_error("eco_pair_fused: C must be <= 64 (got %d); use eco_pair_stats/finalize/grad", C); return -3; }
    if (!upstream || !sums_out || !losses_out) { set_error("null upstream / sums_out / losses_out"); return -5; }
    if (!ws || ws_bytes < eco_pair_fused_ws_bytes(C)) { set_error("workspace too small"); return -5; }
    const bool want_a = ga && ga->ptr, want_b = gb && gb->ptr;
    if (want_a && ga->dtype != a->dtype) { set_error("ga dtype must match slot a"); return -7; }
    if (want_b && gb->dtype != b->dtype) { set_error("gb dtype must match slot b"); return -7; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    bool v4 = aligned_for_vec4(a->ptr, a->sn, a->sc, a->dtype, HW) && aligned_for_vec4(b->ptr, b->sn, b->sc, b->dtype, HW);
    if (want_a) { g.ga = ga->ptr; g.ga_sn = ga->sn; g.ga_sc = ga->sc; v4 = v4 && aligned_for_vec4(ga->ptr, ga->sn, ga->sc, ga->dtype, HW); }
    if (want_b) { g.gb = gb->ptr; g.gb_sn = gb->sn; g.gb_sc = gb->sc; v4 = v4 && aligned_for_vec4(gb->ptr, gb->sn, gb->sc, gb->dtype, HW); }
    const int vec = v4 ? 4 : 1;
    float gamma;
    const bool gen = general_focal(shape_host, gamma);
    dim3 grid;
    rc = plan(g.p, vec, device, grid);
    if (rc) return rc;
    g.p.focal_gamma = gamma;
    const int sms = sm_count_cached(device);
    if ((int64_t)grid.x * grid.y > (int64_t)sms * kCtasPerSm) {
        set_error("eco_pair_fused: %d leaves do not fit one resident wave; use eco_pair_stats/finalize/grad", C);
        return -8;
    }
    fa.bw = background_weight;
    fa.scale = scale;
    if (shape_host && (shape_host->tversky_alpha != 0.5 || shape_host->tversky_beta != 0.3 || shape_host->focal_dice_gamma != 1.8)) {
        fa.shaped = 1;
        fa.shape.alpha = shape_host->tversky_alpha;
        fa.shape.beta = shape_host->tversky_beta;
        fa.shape.fd_gamma = shape_host->focal_dice_gamma;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // one fully contiguous fp32 leaf that fits the GPU's shared memory (cfg1): the resident kernel reads it once
    const bool flat = C == 1 && vec == 4 && !gen && a->dtype == ECO_F32 && b->dtype == ECO_F32 &&
                      (N == 1 || (a->sn == HW && b->sn == HW)) && (!want_a || N == 1 || ga->sn == HW) && (!want_b || N == 1 || gb->sn == HW);
    if (flat) {
        const int64_t total = (int64_t)N * HW;
        int rgrid = sms;
        if ((int64_t)rgrid > total / 4) rgrid = (int)(total / 4);
        const int smem = resident::resident_smem_bytes(total, rgrid);
        if (smem > 0) {
            static thread_local int attr_done[64];
            if (device >= 0 && device < 64 && !attr_done[device]) {
                ECO_CUDA(cudaFuncSetAttribute(resident::leaf_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, resident::kRMaxSmem));
                attr_done[device] = 1;
            }
            resident::ResidentWs* rws = reinterpret_cast<resident::ResidentWs*>(reinterpret_cast<char*>(ws) + resident_ws_offset(C));
            int64_t total_arg = total;
            void* rargs[] = {&fa, &total_arg, (void*)&upstream, &rws, &sums_out, &losses_out};
            return check_cuda(cudaLaunchCooperativeKernel((const void*)resident::leaf_resident_kernel, dim3(rgrid), dim3(resident::kRThreads),
                                                          rargs, smem, st), "leaf_resident_kernel launch");
        }
    }
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + pair_ws_partials_offset(C));
    FusedWs* fw = reinterpret_cast<FusedWs*>(reinterpret_cast<char*>(ws) + fused_ws_offset(C));
    void* args[] = {&fa, (void*)&upstream, &counters, &partials, &fw, &sums_out, &losses_out};
    const void* fn = nullptr;
#define ECO_PICK(TA, TB, V, G) fn = (const void*)pair_fused_kernel<TA, TB, V, G>
#define ECO_PICK_T(V, G)                                                                       \
    do {                                                                                       \
        if (a->dtype == ECO_F32 && b->dtype == ECO_F32) ECO_PICK(float, float, V, G);          \
        else if (a->dtype == ECO_BF16 && b->dtype == ECO_F32) ECO_PICK(__nv_bfloat16, float, V, G); \
        else if (a->dtype == ECO_F32 && b->dtype == ECO_BF16) ECO_PICK(float, __nv_bfloat16, V, G); \
        else ECO_PICK(__nv_bfloat16, __nv_bfloat16, V, G);                                     \
    } while (0)
    if (vec == 4) { if (gen) ECO_PICK_T(4, true); else ECO_PICK_T(4, false); }
    else { if (gen) ECO_PICK_T(1, true); else ECO_PICK_T(1, false); }
#undef ECO_PICK_T
#undef ECO_PICK
    return check_cuda(cudaLaunchCooperativeKernel(fn, grid, dim3(kThreads), args, 0, st), "pair_fused_kernel launch");
}

extern "C" int eco_pair_fused(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                              double background_weight, double scale, const EcoLeafShape* shape_host,
                              const float* upstream, void* ws, int64_t ws_bytes, double* sums_out, float* losses_out,
                              const EcoOut* ga, const EcoOut* gb, int device, void* stream) {
    return eco_pair_fused_ex(a, b, N, C, HW, flags, background_weight, scale, shape_host, upstream, nullptr, ws, ws_bytes, sums_out,
                             losses_out, ga, gb, device, stream);
}

#ifdef ECO_LEAF_TIMELINE
extern "C" int eco_debug_leaf_timeline(unsigned long long* host_out, int n_words) {
    return (int)cudaMemcpyFromSymbol(host_out, g_leaf_tl, (size_t)n_words * 8);
}
#endif
