// Evaluation scoring: sigmoid -> strict '>' threshold -> exact per-class pixel counts, plus the
// un-thresholded soft-Dice sums, in ONE read of logits and labels (8 B/element, HBM-bound).
// Replaces ess/test_multiclass.py:58,68-69,80-81 (see include/ecoloss.h).
#include <type_traits>

#include "eco_common.cuh"

namespace eco {

constexpr int kEvThreads = 256;
constexpr int kEvUnroll = 4;
constexpr int kEvCtasPerSm = 4;
constexpr int kEvMaxCtas = 148 * kEvCtasPerSm * 2;  // per channel

struct EvalArgs {
    const void* z;
    const void* l;
    int64_t z_sn, z_sc, l_sn, l_sc;
    int32_t N, C;
    int64_t HW;
    int32_t tiles_per_plane;
    int64_t tiles_per_channel;
    int32_t tiles_per_cta;
    int32_t n_thr;
    int32_t probs;  // inputs already are probabilities
    int32_t unun;   // soft Dice only: the prediction un-union is applied at load (see dice_counts_kernel)
};

// ws: counters u32[C] (padded to 256 B) | int64 partial counts [C][kEvMaxCtas][NT][2] ... laid out
// generically as: per (c, cta): long long cnt[2*NT + 1] (last = label count), double soft[3].
__host__ __device__ inline int64_t eval_rec_words(int nt) { return 2 * (int64_t)nt + 1 + 3; }  // 8-byte words
__host__ __device__ inline int64_t eval_ws_off(int C) { return ((int64_t)C * 8 + 255) / 256 * 256; }   // counters u32[C] | non-binary flags u32[C]

// Labels other than exactly 0 / 1 (the datasets that resize their masks produce fractions, ess/dataset/fish/fish_suim.py:60-74):
// the reference's thresholded Dice is 2 sum(out * lab) / (sum out + sum lab^2) with the REAL label values
// (ess/test_multiclass.py:80 -> ess/loss_functions.py:55-57), which integer counts cannot express.  The counting kernels
// count a label only where it is exactly 1, so sum lab^2 == count holds iff every label of the class is 0 or 1 (all terms
// are non-negative: no cancellation).  The last CTA of a class checks that, publishes the per-threshold intersections as
// float64 (= the integer counts for binary labels) and raises the class's flag otherwise; dice_exact_kernel, launched
// after every thresholded scoring call, returns at once unless a flag is up and then re-reads the class for the exact sums.
__device__ __forceinline__ void publish_intersections(int c, int C, int n_thr, long long* counts_out, const double* soft_out,
                                                      double* thr_inter_out, unsigned int* flags) {
    if (n_thr <= 0) { flags[c] = 0u; return; }
    const long long lab_cnt = counts_out[(int64_t)c * 3 + 2];
    const bool nb = soft_out[c * 3 + 2] != (double)lab_cnt;
    flags[c] = (nb && thr_inter_out) ? 1u : 0u;
    for (int k = 0; k < n_thr; ++k) {
        long long* r = counts_out + ((int64_t)k * C + c) * 3;
        if (thr_inter_out) thr_inter_out[(int64_t)k * C + c] = nb ? 0.0 : (double)r[0];
        else if (nb) r[2] = -1;   // no place for the exact sums: mark the class, eco_dice_finalize turns it into NaN
    }
}

// Per-thread counters are packed: low 16 bits = |out| count, high 16 bits = intersection count of one threshold
// (one FSETP + one predicated IADD per threshold and element); a thread folds them into 32-bit counters before
// either half can overflow.  The sigmoid is the 4-instruction MUFU form and the exact (ATen-bit-compatible) one is
// recomputed only within 4e-6 of a threshold, which keeps the counts bit-exact.  (0 or 1 threshold; more go to the beam.)
constexpr int kPackFlushElems = 32768;
constexpr float kThrEps = 4e-6f;

//
// UNUN (soft Dice only, NT == 0): the prediction un-union of the sequential model's test
// (ess/test_multiclass_sequential_densenetloss.py:66 -> ess/utils/subsets_union.py:21-27, reverse=True with
// exclude_indices=[0]: for idx = C-2 .. 1, out[:, idx] = |out[:, idx] - out[:, idx + 1]| in place, i.e. the recursion
// p'_c = |p_c - p'_{c+1}|, p'_{C-1} = p_{C-1}) is taken in registers at load: a CTA of channel c in 1 .. C-2 also reads
// the planes c+1 .. C-1 of its elements (L2 hits: the CTAs of those channels stream the same lines) instead of a
// separate in-place sweep over the predictions.  Channels 0 and C-1 run the plain code.
template <typename TZ, typename TL, int VEC, int NT, bool UNUN = false>
__global__ void __launch_bounds__(kEvThreads, UNUN ? kEvCtasPerSm - 1 : kEvCtasPerSm)   // (the running p' costs 16 registers)
dice_counts_kernel(EvalArgs p, const float* __restrict__ thresholds, unsigned int* __restrict__ counters,
                   long long* __restrict__ partials, long long* __restrict__ counts_out,
                   double* __restrict__ soft_out, double* __restrict__ thr_inter_out) {
    constexpr int kTile = kEvThreads * VEC * kEvUnroll;
    constexpr int NTA = NT > 0 ? NT : 1;
    const int c = blockIdx.y;
    const TZ* __restrict__ zbase = reinterpret_cast<const TZ*>(p.z) + (int64_t)c * p.z_sc;
    const TL* __restrict__ lbase = reinterpret_cast<const TL*>(p.l) + (int64_t)c * p.l_sc;
    static_assert(!UNUN || NT == 0, "the un-union is fused for the soft Dice only");
    const bool uu = UNUN && c >= 1 && c <= p.C - 2;

    float thr[NTA];
#pragma unroll
    for (int k = 0; k < NTA; ++k) thr[k] = (k < p.n_thr) ? thresholds[k] : __int_as_float(0x7f800000);

    // per-thread exact counters (a thread sees far fewer than 2^31 elements)
    unsigned int cnt_pk[NTA];
    int cnt_l = 0;
    int since_fold = 0;
#pragma unroll
    for (int k = 0; k < NTA; ++k) cnt_pk[k] = 0u;
    __shared__ long long sm_cnt[2 * NTA + 1];
    if (threadIdx.x < 2 * NTA + 1) sm_cnt[threadIdx.x] = 0;
    __syncthreads();
    double dsoft[3] = {0.0, 0.0, 0.0};

    int64_t tile = (int64_t)blockIdx.x * p.tiles_per_cta;
    int64_t tile_end = tile + p.tiles_per_cta;
    if (tile_end > p.tiles_per_channel) tile_end = p.tiles_per_channel;
    int64_t n = tile / p.tiles_per_plane;
    int32_t t = (int32_t)(tile - n * p.tiles_per_plane);

    // One tile.  FULL: the tile lies inside its plane (every tile but a plane's last): no bounds checks, and the four loads per
    // operand are one base pointer + immediate offsets -- with the checks the compiler re-derives the 64-bit image offset for
    // every load, 8 of the 26-29 instructions per element.  Only the byte-label instantiations get the second copy of the
    // loop: they are bound by their instruction stream (5 B/element), the others by HBM, and with 16 label registers more the
    // duplicated loop spills under the 64-register budget.
    auto tile_body = [&](auto full_c) {
        constexpr bool FULL = decltype(full_c)::value;
        const int64_t e0 = (int64_t)t * kTile + (int64_t)threadIdx.x * VEC;
        const TZ* zp = zbase + n * p.z_sn + e0;
        const TL* lp = lbase + n * p.l_sn + e0;
        float zv[kEvUnroll][VEC], lv[kEvUnroll][VEC];
        bool ok[kEvUnroll];
#pragma unroll
        for (int u = 0; u < kEvUnroll; ++u) {
            ok[u] = FULL || e0 + (int64_t)u * kEvThreads * VEC < p.HW;
            if (ok[u]) {
                if constexpr (VEC == 4) {
                    Vec4<TZ>::load(zp + u * kEvThreads * VEC, reinterpret_cast<float(&)[4]>(zv[u]));
                    Vec4<TL>::load(lp + u * kEvThreads * VEC, reinterpret_cast<float(&)[4]>(lv[u]));
                } else {
                    zv[u][0] = Vec4<TZ>::load1(zp + u * kEvThreads * VEC);
                    lv[u][0] = Vec4<TL>::load1(lp + u * kEvThreads * VEC);
                }
            }
        }
        float below[UNUN ? kEvUnroll : 1][VEC];   // p'_{c+1} of this thread's elements
        if (UNUN && uu) {
            for (int k = p.C - 1; k > c; --k) {
                const TZ* zk = zp + (int64_t)(k - c) * p.z_sc;
#pragma unroll
                for (int u = 0; u < kEvUnroll; ++u) {
                    if (!ok[u]) continue;
                    float t4[VEC];
                    if constexpr (VEC == 4) Vec4<TZ>::load(zk + u * kEvThreads * VEC, reinterpret_cast<float(&)[4]>(t4));
                    else t4[0] = Vec4<TZ>::load1(zk + u * kEvThreads * VEC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const float pk = p.probs ? t4[v] : sigmoid_fast(t4[v]);
                        below[UNUN ? u : 0][v] = (k == p.C - 1) ? pk : fabsf(pk - below[UNUN ? u : 0][v]);
                    }
                }
            }
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < kEvUnroll; ++u) {
            if (!ok[u]) continue;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float pr;
                if (p.probs) {
                    pr = zv[u][v];
                } else {
                    pr = sigmoid_fast(zv[u][v]);
                    if (NT > 0) {
                        bool near = false;
#pragma unroll
                        for (int k = 0; k < NTA; ++k) near |= fabsf(pr - thr[k]) < kThrEps;
                        if (near) pr = sigmoid_exact(zv[u][v]);  // rare: the strict '>' must see ATen's bits
                    }
                }
                if (UNUN && uu) pr = fabsf(pr - below[UNUN ? u : 0][v]);
                const float lab = lv[u][v];
                const int li = lab == 1.0f ? 1 : 0;
                const unsigned int inc = 1u + ((unsigned int)li << 16);
                s0 = fmaf(pr, lab, s0);
                s1 += pr;
                s2 = fmaf(lab, lab, s2);
                cnt_l += li;
                if (NT > 0) {
#pragma unroll
                    for (int k = 0; k < NTA; ++k) cnt_pk[k] += (pr > thr[k]) ? inc : 0u;
                }
            }
        }
        dsoft[0] += (double)s0;
        dsoft[1] += (double)s1;
        dsoft[2] += (double)s2;
    };

    for (; tile < tile_end; ++tile) {
        if (sizeof(TL) == 1 && !UNUN && (int64_t)(t + 1) * kTile <= p.HW) tile_body(std::true_type{});
        else tile_body(std::false_type{});
        if (NT > 0 && (since_fold += kEvUnroll * VEC) >= kPackFlushElems) {  // rare: before a 16-bit half can overflow
#pragma unroll
            for (int k = 0; k < NTA; ++k) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm_cnt[2 * k]), (unsigned long long)(cnt_pk[k] & 0xffffu));
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm_cnt[2 * k + 1]), (unsigned long long)(cnt_pk[k] >> 16));
                cnt_pk[k] = 0u;
            }
            since_fold = 0;
        }
        if (++t == p.tiles_per_plane) {
            t = 0;
            ++n;
        }
    }

    // CTA reduction: integers via warp shuffles + shared atomics (exact, order-independent)
    __shared__ double sm_soft[kEvThreads / 32][3];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (NT > 0) {
#pragma unroll
        for (int k = 0; k < NTA; ++k) {
            int o = __reduce_add_sync(0xffffffffu, (int)(cnt_pk[k] & 0xffffu));
            int i = __reduce_add_sync(0xffffffffu, (int)(cnt_pk[k] >> 16));
            if (lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm_cnt[2 * k]), (unsigned long long)o);
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm_cnt[2 * k + 1]), (unsigned long long)i);
            }
        }
    }
    {
        int l = __reduce_add_sync(0xffffffffu, cnt_l);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&sm_cnt[2 * NTA]), (unsigned long long)l);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = warp_sum(dsoft[k]);
        if (lane == 0) sm_soft[warp][k] = v;
    }
    __syncthreads();
    const int64_t rec = eval_rec_words(NTA);
    long long* mine = partials + ((int64_t)c * kEvMaxCtas + blockIdx.x) * rec;
    if (threadIdx.x < 2 * NTA + 1) mine[threadIdx.x] = sm_cnt[threadIdx.x];
    if (threadIdx.x < 3) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kEvThreads / 32; ++w) v += sm_soft[w][threadIdx.x];
        reinterpret_cast<double*>(mine)[2 * NTA + 1 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(&counters[c], 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const long long* base = partials + (int64_t)c * kEvMaxCtas * rec;
    // counts: one warp per record word, lanes stride over CTAs
    for (int w = warp; w < 2 * NTA + 1 + 3; w += kEvThreads / 32) {
        if (w < 2 * NTA + 1) {
            long long v = 0;
            for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(base + (int64_t)i * rec + w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                if (w == 2 * NTA) {
                    // label count goes to every threshold's third slot
                    for (int k = 0; k < (p.n_thr > 0 ? p.n_thr : 1); ++k)
                        counts_out[((int64_t)k * p.C + c) * 3 + 2] = v;
                } else {
                    const int k = w >> 1;
                    // order in counts_out: (intersection, |out|, |lab|)
                    if (k < p.n_thr) counts_out[((int64_t)k * p.C + c) * 3 + ((w & 1) ? 0 : 1)] = v;
                }
            }
        } else {
            const int s = w - (2 * NTA + 1);
            double v = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32)
                v += __ldcg(reinterpret_cast<const double*>(base + (int64_t)i * rec) + 2 * NTA + 1 + s);
            v = warp_sum(v);
            if (lane == 0) soft_out[c * 3 + s] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        publish_intersections(c, p.C, p.n_thr, counts_out, soft_out, thr_inter_out, counters + p.C);
        counters[c] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Threshold beam (2..20 thresholds, ess/test_multiclass.py:64-77) by BINNING instead of one compare per threshold:
// the thresholds are sorted once per CTA; an element's bin = number of thresholds below its probability and it bumps
// ONE packed counter {low 16 bits: elements, high 16 bits: elements with label 1} of its thread-private
// shared-memory histogram (a plain read-modify-write; a shared-memory atomic costs 2 cycles per lane).
// counts(T_r) = sum of the bins above r.  The bin comes from a 256-cell direct-lookup table when the thresholds' cells
// are at least three apart (the reference's np.arange(0.8, 0.99, 0.01)): beam_group_lut, ~19 instructions per element,
// HBM-bound (19 thresholds at cfg3: 227 us = 93 % of the measured peak; profiles/r2b_dice_beam_ncu_full.json).  Otherwise
// a 5-step branch-free search (two levels in registers, three in shared memory) over the whole tile first:
// beam_tile_search, the round-1 form of the kernel (issue-bound at ~57 instructions per element, 374 us).  Either way the
// sigmoid is the 4-instruction MUFU form and ATen's exact bits are recomputed only for elements within kThrEps of a
// threshold.
// ---------------------------------------------------------------------------------------------
constexpr int kBeamMax = 20;
constexpr int kBeamBins = kBeamMax + 1;
constexpr int kBeamThreads = 256;
constexpr int kBeamCells = 256;   // power of two
#ifndef ECO_BEAM_CTAS
#define ECO_BEAM_CTAS 3
#endif
constexpr int kBeamCtasPerSm = ECO_BEAM_CTAS;   // 80 registers (64 spill in the streaming loop); measured 2 / 3 / 4 CTAs per SM within 3 %

struct BeamSmem {
    float sorted[32];            // [0..30] ascending thresholds padded with +inf (31 entries are searched)
    float2 nbr[32];              // per bin b: {largest threshold below, smallest threshold at or above}
    int rank_of[kBeamMax];       // original index k -> position in the sorted order
    unsigned int hist[kBeamBins][kBeamThreads];
    long long cnt[2 * kBeamMax + 1];
    double soft[kBeamThreads / 32][3];
    bool is_last;
    // direct lookup (finite, distinct thresholds whose cells are at least three apart, e.g. the reference's
    // np.arange(0.8, 0.99, 0.01)): cell j of a uniform grid over [T_min, T_max] -> {byte offset of histogram row `rank`,
    // T} where T is THE threshold whose cell is j - 1, j or j + 1 and `rank` the number of thresholds below it; without
    // such a threshold {row of the thresholds in the cells below, +inf}.  bin = rank + (p > T) either way.
    uint2 lut[kBeamCells];
    float lut_scale, lut_bias;
    int lut_ok;
    uint32_t lut_adj;            // shared-memory address of lut[] minus (kCellBits0 << 3): entry = lut_adj + (cell bits << 3)
    int cell_of[32];
};

// cell of a value, as the bits 0x4b000000 + cell: ONE monotone function for thresholds and probabilities alike, so every
// threshold in a lower cell is below the value and every threshold in a higher cell above it, whatever the rounding of the
// two FMAs.  The saturating FMA clamps (and sends NaN to cell 0); no F2I: adding 2^23 leaves the integer in the mantissa.
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
constexpr unsigned int kCellBits0 = 0x4b000000u;
__device__ __forceinline__ unsigned int beam_cell_bits(float v, float scale, float bias) {
    return __float_as_uint(fmaf(fma_sat(v, scale, bias), (float)(kBeamCells - 1), 8388608.0f));
}
__device__ __forceinline__ int beam_cell(float v, float scale, float bias) { return (int)(beam_cell_bits(v, scale, bias) - kCellBits0); }
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
constexpr unsigned int kBeamRow = kBeamThreads * 4;   // bytes per histogram row

__device__ __forceinline__ int beam_bin(float p, const BeamSmem& sm, float t15, float t7, float t23) {
    int b = p > t15 ? 16 : 0;
    b += (p > (b ? t23 : t7)) ? 8 : 0;
    b += (p > sm.sorted[b + 3]) ? 4 : 0;
    b += (p > sm.sorted[b + 1]) ? 2 : 0;
    b += (p > sm.sorted[b]) ? 1 : 0;
    return b;   // number of thresholds T with p > T (strict, fp32), 0..n_thr
}

// One vector of VEC elements on the direct-lookup path: probability -> cell -> {row, T} -> row + (p > T) -> ONE packed
// counter of the thread's histogram column (a plain LDS / IADD / STS: the column is private).  An element within kThrEps
// of T is doubtful (the approximate sigmoid may sit on the other side of T than ATen's): then -- rarely -- the vector's
// rows are redone from the exact sigmoid.  The soft sums of two elements share packed fp32x2 instructions.
// ~19 instructions per element all told (the first version of the beam: 57).
__device__ __forceinline__ void hist_bump(uint32_t addr, unsigned int inc) {
    unsigned int h;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(h) : "r"(addr));
    h += inc;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(h) : "memory");
}
__device__ __forceinline__ float2 sigmoid_fast_pair(float2 z) {
    const float2 t = __fmul2_rn(z, make_float2(-kLog2e, -kLog2e));
    const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.0f, 1.0f));
    return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}
struct BeamSoft {
    float2 pl, p, ll;   // sum p * lab, sum p, sum lab^2 (two lanes each)
};
template <int VEC, bool PROBS>
__device__ __forceinline__ void beam_group_lut(const float (&z)[VEC], const float (&lab)[VEC], uint32_t lut_adj, float scale,
                                               float bias, uint32_t hist_tid, BeamSoft& acc) {
    float pr[VEC];
    if constexpr (VEC == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 zz = make_float2(z[2 * h], z[2 * h + 1]), ll = make_float2(lab[2 * h], lab[2 * h + 1]);
            const float2 pp = PROBS ? zz : sigmoid_fast_pair(zz);
            pr[2 * h] = pp.x; pr[2 * h + 1] = pp.y;
            acc.pl = __ffma2_rn(pp, ll, acc.pl);
            acc.p = __fadd2_rn(acc.p, pp);
            acc.ll = __ffma2_rn(ll, ll, acc.ll);
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            pr[v] = PROBS ? z[v] : sigmoid_fast(z[v]);
            acc.pl.x = fmaf(pr[v], lab[v], acc.pl.x); acc.p.x += pr[v]; acc.ll.x = fmaf(lab[v], lab[v], acc.ll.x);
        }
    }
    uint32_t addr[VEC];   // this thread's counter of the element's bin
    bool near = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const uint2 e = lds_u2(lut_adj + (beam_cell_bits(pr[v], scale, bias) << 3));
        // p > T  <=>  T - p is negative: its sign bit (T - p is +0 for p == T, +inf for T = +inf, and a NaN comes out
        // canonical, i.e. positive: a NaN is never above a threshold, as in the reference's out[out > T] = 1)
        const float d = __uint_as_float(e.y) - pr[v];
        addr[v] = hist_tid + e.x + ((__float_as_uint(d) >> 21) & kBeamRow);
        near = near || fabsf(d) < kThrEps;
    }
    if (!PROBS && near) {   // rare: the strict '>' must see ATen's bits
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float pe = sigmoid_exact(z[v]);
            const uint2 e = lds_u2(lut_adj + (beam_cell_bits(pe, scale, bias) << 3));
            addr[v] = hist_tid + e.x + (pe > __uint_as_float(e.y) ? kBeamRow : 0u);
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) hist_bump(addr[v], lab[v] == 1.0f ? 0x10001u : 1u);
}
static_assert(kBeamRow == 1u << 10, "the row select above takes the sign bit down to bit 10");

// search path (any thresholds): the whole tile branch-free first (the MUFU / LDS latencies of its elements overlap), then
// the rare exact redo, then the histogram
template <int VEC>
__device__ __forceinline__ void beam_tile_search(const float (&zv)[kEvUnroll][VEC], const float (&lv)[kEvUnroll][VEC],
                                                 const bool (&ok)[kEvUnroll], bool probs, BeamSmem& sm, float t15, float t7,
                                                 float t23, int tid, BeamSoft& acc) {
    int bins[kEvUnroll][VEC];
    unsigned int redo = 0u;
#pragma unroll
    for (int u = 0; u < kEvUnroll; ++u) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float pr = probs ? zv[u][v] : sigmoid_fast(zv[u][v]);
            const int b = beam_bin(pr, sm, t15, t7, t23);
            bins[u][v] = b;
            const float2 nb = sm.nbr[b];
            const bool near = !probs && ((pr - nb.x) < kThrEps || (nb.y - pr) < kThrEps);
            redo |= near ? (1u << (u * VEC + v)) : 0u;
            const float lab = lv[u][v];
            acc.pl.x = fmaf(pr, lab, acc.pl.x);
            acc.p.x += pr;
            acc.ll.x = fmaf(lab, lab, acc.ll.x);
        }
    }
    if (redo) {   // rare: the strict '>' must see ATen's bits
#pragma unroll
        for (int u = 0; u < kEvUnroll; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (redo & (1u << (u * VEC + v))) bins[u][v] = beam_bin(sigmoid_exact(zv[u][v]), sm, t15, t7, t23);
    }
#pragma unroll
    for (int u = 0; u < kEvUnroll; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const unsigned int li = lv[u][v] == 1.0f ? 0x10000u : 0u;
            sm.hist[bins[u][v]][tid] += 1u + li;   // thread-private slot: plain LDS / IADD / STS
        }
    }
}

// the streaming loop of one CTA; MODE 0 = search path, 1 = direct lookup from logits, 2 = direct lookup from probabilities
template <typename TZ, typename TL, int VEC, int MODE>
__device__ __forceinline__ void beam_stream(const EvalArgs& p, const TZ* __restrict__ zbase, const TL* __restrict__ lbase,
                                            BeamSmem& sm, int tid, double (&dsoft)[3]) {
    constexpr int kTile = kBeamThreads * VEC * kEvUnroll;
    constexpr int NTA = kBeamMax;
    const float lut_scale = sm.lut_scale, lut_bias = sm.lut_bias;
    const float t15 = sm.sorted[15], t7 = sm.sorted[7], t23 = sm.sorted[23];
    // shared-memory addresses with the constant parts folded in: table entry = lut_adj + (cell bits << 3) (read back from
    // shared memory: as a compile-time constant ptxas splits it into two adds per element), histogram column = hist_tid +
    // row offset
    const uint32_t lut_adj = sm.lut_adj;
    const uint32_t hist_tid = (uint32_t)__cvta_generic_to_shared(&sm.hist[0][tid]);
    int since_fold = 0;
    int64_t tile = (int64_t)blockIdx.x * p.tiles_per_cta;
    int64_t tile_end = tile + p.tiles_per_cta;
    if (tile_end > p.tiles_per_channel) tile_end = p.tiles_per_channel;
    int64_t n = tile / p.tiles_per_plane;
    int32_t t = (int32_t)(tile - n * p.tiles_per_plane);

    for (; tile < tile_end; ++tile) {
        const TZ* zp = zbase + n * p.z_sn;
        const TL* lp = lbase + n * p.l_sn;
        const int64_t e0 = (int64_t)t * kTile + (int64_t)tid * VEC;
        float zv[kEvUnroll][VEC], lv[kEvUnroll][VEC];
        BeamSoft acc;
        acc.pl = acc.p = acc.ll = make_float2(0.f, 0.f);
        if (MODE != 0 && (int64_t)(t + 1) * kTile <= p.HW) {
            // the whole tile lies inside the plane (all but the last tile of a plane): no per-vector bounds checks
#pragma unroll
            for (int u = 0; u < kEvUnroll; ++u) {
                const int64_t e = e0 + (int64_t)u * kBeamThreads * VEC;
                if constexpr (VEC == 4) {
                    Vec4<TZ>::load(zp + e, reinterpret_cast<float(&)[4]>(zv[u]));
                    Vec4<TL>::load(lp + e, reinterpret_cast<float(&)[4]>(lv[u]));
                } else {
                    zv[u][0] = Vec4<TZ>::load1(zp + e);
                    lv[u][0] = Vec4<TL>::load1(lp + e);
                }
            }
#pragma unroll
            for (int u = 0; u < kEvUnroll; ++u)
                beam_group_lut<VEC, MODE == 2>(zv[u], lv[u], lut_adj, lut_scale, lut_bias, hist_tid, acc);
        } else {
        bool ok[kEvUnroll];
#pragma unroll
        for (int u = 0; u < kEvUnroll; ++u) {
            const int64_t e = e0 + (int64_t)u * kBeamThreads * VEC;
            ok[u] = e < p.HW;
            if (ok[u]) {
                if constexpr (VEC == 4) {
                    Vec4<TZ>::load(zp + e, reinterpret_cast<float(&)[4]>(zv[u]));
                    Vec4<TL>::load(lp + e, reinterpret_cast<float(&)[4]>(lv[u]));
                } else {
                    zv[u][0] = Vec4<TZ>::load1(zp + e);
                    lv[u][0] = Vec4<TL>::load1(lp + e);
                }
            } else if (MODE == 0) {
                // outside the plane: probability 0 and label 0 add nothing to the soft sums; the bins are not counted
#pragma unroll
                for (int v = 0; v < VEC; ++v) { zv[u][v] = p.probs ? 0.f : -__int_as_float(0x7f800000); lv[u][v] = 0.f; }
            }
        }
        if constexpr (MODE == 0) {
            beam_tile_search<VEC>(zv, lv, ok, p.probs != 0, sm, t15, t7, t23, tid, acc);
        } else {
#pragma unroll
            for (int u = 0; u < kEvUnroll; ++u)
                if (ok[u]) beam_group_lut<VEC, MODE == 2>(zv[u], lv[u], lut_adj, lut_scale, lut_bias, hist_tid, acc);
        }
        }
        dsoft[0] += (double)(acc.pl.x + acc.pl.y);
        dsoft[1] += (double)(acc.p.x + acc.p.y);
        dsoft[2] += (double)(acc.ll.x + acc.ll.y);
        if ((since_fold += kEvUnroll * VEC) >= kPackFlushElems) {   // before a 16-bit half can overflow (huge inputs only)
#pragma unroll 1
            for (int b = 0; b < kBeamBins; ++b) {
                const unsigned int h = sm.hist[b][tid];
                sm.hist[b][tid] = 0u;
                if (h == 0u) continue;
                // bin b counts towards the sorted thresholds r < b; every bin counts towards the label total
                for (int r = 0; r < b && r < p.n_thr; ++r) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * r]), (unsigned long long)(h & 0xffffu));
                    atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * r + 1]), (unsigned long long)(h >> 16));
                }
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * NTA]), (unsigned long long)(h >> 16));
            }
            since_fold = 0;
        }
        if (++t == p.tiles_per_plane) {
            t = 0;
            ++n;
        }
    }
}

template <typename TZ, typename TL, int VEC>
__global__ void __launch_bounds__(kBeamThreads, kBeamCtasPerSm)
dice_beam_kernel(EvalArgs p, const float* __restrict__ thresholds, unsigned int* __restrict__ counters,
                 long long* __restrict__ partials, long long* __restrict__ counts_out, double* __restrict__ soft_out, double* __restrict__ thr_inter_out) {
    constexpr int NTA = kBeamMax;
    __shared__ BeamSmem sm;
    const int c = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TZ* __restrict__ zbase = reinterpret_cast<const TZ*>(p.z) + (int64_t)c * p.z_sc;
    const TL* __restrict__ lbase = reinterpret_cast<const TL*>(p.l) + (int64_t)c * p.l_sc;

    // ---- sort the thresholds (rank = number of smaller ones; ties by index), neighbours per bin ------------------------
    if (tid < 32) sm.sorted[tid] = __int_as_float(0x7f800000);
    if (tid < 2 * NTA + 1) sm.cnt[tid] = 0;
#pragma unroll
    for (int b = 0; b < kBeamBins; ++b) sm.hist[b][tid] = 0u;
    __syncthreads();
    if (tid < p.n_thr) {
        float t = thresholds[tid];
        if (t != t) t = __int_as_float(0x7f800000);   // p > NaN is never true: same as +inf
        int r = 0;
        for (int j = 0; j < p.n_thr; ++j) {
            float u = thresholds[j];
            if (u != u) u = __int_as_float(0x7f800000);
            r += (u < t || (u == t && j < tid)) ? 1 : 0;
        }
        sm.sorted[r] = t;
        sm.rank_of[tid] = r;
    }
    __syncthreads();
    if (tid < 32) sm.nbr[tid] = make_float2(tid > 0 ? sm.sorted[tid - 1] : -__int_as_float(0x7f800000), sm.sorted[tid < 31 ? tid : 31]);
    // ---- direct-lookup table, when the thresholds allow it ----------------------------------------------------------------
    if (tid == 0) {
        const int nt = p.n_thr;
        const float lo = sm.sorted[0], hi = sm.sorted[nt - 1];
        // finite and spread over more than 4 kThrEps per cell: an element within kThrEps of a threshold then sits in the
        // threshold's cell or next to it.  [lo, hi] maps to the cells 1 .. 254.
        const bool ok = nt >= 2 && fabsf(lo) < 1e30f && fabsf(hi) < 1e30f && (hi - lo) > 4.0f * kThrEps * (float)(kBeamCells - 3);
        const float sc = ok ? ((float)(kBeamCells - 3) / (float)(kBeamCells - 1)) / (hi - lo) : 0.0f;
        sm.lut_ok = ok ? 1 : 0;
        sm.lut_scale = sc;
        sm.lut_bias = ok ? 1.0f / (float)(kBeamCells - 1) - lo * sc : 0.0f;
        sm.lut_adj = (uint32_t)__cvta_generic_to_shared(&sm.lut[0]) - (kCellBits0 << 3);
    }
    __syncthreads();
    if (tid < 32) sm.cell_of[tid] = tid < p.n_thr ? beam_cell(sm.sorted[tid], sm.lut_scale, sm.lut_bias) : (1 << 20) + 8 * tid;
    __syncthreads();
    if (tid + 1 < p.n_thr && sm.cell_of[tid + 1] - sm.cell_of[tid] < 3) sm.lut_ok = 0;   // (also: duplicate thresholds)
    __syncthreads();
    if (sm.lut_ok) {
        int below = 0, own = -1;
        for (int r = 0; r < p.n_thr; ++r) {
            const int cr = sm.cell_of[r];
            below += cr < tid ? 1 : 0;
            if (cr >= tid - 1 && cr <= tid + 1) own = r;   // at most one: the cells are three apart
        }
        // (sorted, distinct: the rank of sorted[own] is own)
        sm.lut[tid] = own >= 0 ? make_uint2((unsigned)own * kBeamRow, __float_as_uint(sm.sorted[own]))
                               : make_uint2((unsigned)below * kBeamRow, 0x7f800000u);
    }
    __syncthreads();
    double dsoft[3] = {0.0, 0.0, 0.0};
    if (!sm.lut_ok) beam_stream<TZ, TL, VEC, 0>(p, zbase, lbase, sm, tid, dsoft);
    else if (!p.probs) beam_stream<TZ, TL, VEC, 1>(p, zbase, lbase, sm, tid, dsoft);
    else beam_stream<TZ, TL, VEC, 2>(p, zbase, lbase, sm, tid, dsoft);

    // ---- CTA reduction: per bin over the threads, then suffix sums -> per sorted threshold ----------------------------
    __syncthreads();
    for (int b = warp; b < kBeamBins; b += kBeamThreads / 32) {
        int o = 0, i = 0;
        for (int k = lane; k < kBeamThreads; k += 32) {
            const unsigned int h = sm.hist[b][k];
            o += (int)(h & 0xffffu);
            i += (int)(h >> 16);
        }
        o = __reduce_add_sync(0xffffffffu, o);
        i = __reduce_add_sync(0xffffffffu, i);
        if (lane == 0) {
            for (int r = 0; r < b && r < p.n_thr; ++r) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * r]), (unsigned long long)o);
                atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * r + 1]), (unsigned long long)i);
            }
            atomicAdd(reinterpret_cast<unsigned long long*>(&sm.cnt[2 * NTA]), (unsigned long long)i);   // label total: all bins
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = warp_sum(dsoft[k]);
        if (lane == 0) sm.soft[warp][k] = v;
    }
    __syncthreads();
    // partial record in the layout of dice_counts_kernel<.., 20>: word 2k = |out| and 2k+1 = intersection of the
    // caller's k-th threshold (= sorted position rank_of[k]), word 40 = label count, then the three soft sums
    const int64_t rec = eval_rec_words(NTA);
    long long* mine = partials + ((int64_t)c * kEvMaxCtas + blockIdx.x) * rec;
    if (tid < 2 * NTA) {
        const int k = tid >> 1;
        mine[tid] = k < p.n_thr ? sm.cnt[2 * sm.rank_of[k] + (tid & 1)] : 0;
    }
    if (tid == 2 * NTA) mine[tid] = sm.cnt[2 * NTA];
    if (tid < 3) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kBeamThreads / 32; ++w) v += sm.soft[w][tid];
        reinterpret_cast<double*>(mine)[2 * NTA + 1 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int prev = atomicAdd(&counters[c], 1u);
        sm.is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!sm.is_last) return;
    __threadfence();
    const long long* base = partials + (int64_t)c * kEvMaxCtas * rec;
    for (int w = warp; w < 2 * NTA + 1 + 3; w += kBeamThreads / 32) {
        if (w < 2 * NTA + 1) {
            long long v = 0;
            for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(base + (int64_t)i * rec + w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                if (w == 2 * NTA) {
                    for (int k = 0; k < p.n_thr; ++k) counts_out[((int64_t)k * p.C + c) * 3 + 2] = v;
                } else {
                    const int k = w >> 1;
                    if (k < p.n_thr) counts_out[((int64_t)k * p.C + c) * 3 + ((w & 1) ? 0 : 1)] = v;
                }
            }
        } else {
            const int sidx = w - (2 * NTA + 1);
            double v = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32)
                v += __ldcg(reinterpret_cast<const double*>(base + (int64_t)i * rec) + 2 * NTA + 1 + sidx);
            v = warp_sum(v);
            if (lane == 0) soft_out[c * 3 + sidx] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        publish_intersections(c, p.C, p.n_thr, counts_out, soft_out, thr_inter_out, counters + p.C);
        counters[c] = 0;
    }
}

// Rare exact pass for classes whose labels are not all 0 / 1 (see publish_intersections): sum out * lab per threshold.
template <typename TZ, typename TL>
__global__ void __launch_bounds__(256)
dice_exact_kernel(EvalArgs p, const float* __restrict__ thresholds, const unsigned int* __restrict__ flags,
                  double* __restrict__ thr_inter_out) {
    const int c = blockIdx.y;
    if (!flags[c]) return;
    __shared__ float thr[kBeamMax];
    __shared__ double red[8][kBeamMax];
    if (threadIdx.x < kBeamMax) thr[threadIdx.x] = threadIdx.x < p.n_thr ? thresholds[threadIdx.x] : __int_as_float(0x7f800000);
    __syncthreads();
    const TZ* zbase = reinterpret_cast<const TZ*>(p.z) + (int64_t)c * p.z_sc;
    const TL* lbase = reinterpret_cast<const TL*>(p.l) + (int64_t)c * p.l_sc;
    double acc[kBeamMax];
#pragma unroll
    for (int k = 0; k < kBeamMax; ++k) acc[k] = 0.0;
    const int64_t total = (int64_t)p.N * p.HW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = e / p.HW, i = e - n * p.HW;
        const float zv = Vec4<TZ>::load1(zbase + n * p.z_sn + i);
        const float lab = Vec4<TL>::load1(lbase + n * p.l_sn + i);
        const float pr = p.probs ? zv : sigmoid_exact(zv);
#pragma unroll
        for (int k = 0; k < kBeamMax; ++k)
            if (pr > thr[k]) acc[k] += (double)lab;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kBeamMax; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < p.n_thr) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        atomicAdd(&thr_inter_out[(int64_t)threadIdx.x * p.C + c], v);
    }
}

struct DiceFinArgs {
    int C, n_thr;
};
__global__ void dice_finalize_kernel(const long long* __restrict__ counts, const double* __restrict__ soft,
                                     const double* __restrict__ thr_inter, DiceFinArgs a, float* __restrict__ dice_out,
                                     float* __restrict__ soft_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double eps = 1e-7;
    if (dice_out && counts && i < a.n_thr * a.C) {
        const long long* r = counts + (int64_t)i * 3;
        if (thr_inter && soft) {
            // the reference's formula with the real label values: 2 sum(out*lab) / (sum out + sum lab^2); for 0/1 labels
            // these are the integer counts
            dice_out[i] = (float)((2.0 * thr_inter[i] + eps) / ((double)r[1] + soft[(int64_t)(i % a.C) * 3 + 2] + eps));
        } else if (r[2] < 0) {
            dice_out[i] = __int_as_float(0x7fc00000);   // labels other than 0/1 and no exact sums were requested
        } else {
            dice_out[i] = (float)((2.0 * (double)r[0] + eps) / ((double)r[1] + (double)r[2] + eps));
        }
    }
    if (soft_out && soft && i < a.C) {
        const double* r = soft + (int64_t)i * 3;
        soft_out[i] = (float)((2.0 * r[0] + eps) / ((r[1] + r[2]) + eps));
    }
}

static bool ev_aligned(const EcoView* v, int64_t HW) {
    const int64_t esz = v->dtype == ECO_BF16 ? 2 : (v->dtype == ECO_U8 ? 1 : 4);
    return (reinterpret_cast<uintptr_t>(v->ptr) % (4 * esz) == 0) && (v->sn % 4 == 0) && (v->sc % 4 == 0) && (HW % 4 == 0);
}

template <int NT>
static void launch_nt(const EvalArgs& p, int zd, int ld, int vec, dim3 grid, cudaStream_t st, const float* thr,
                      unsigned int* counters, long long* partials, long long* counts_out, double* soft_out, double* inter) {
#define ECO_EV(TZ, TL, V)                                                                                                       \
    do {                                                                                                                        \
        if constexpr (NT == 0) {                                                                                                \
            if (p.unun) { dice_counts_kernel<TZ, TL, V, 0, true><<<grid, kEvThreads, 0, st>>>(p, thr, counters, partials, counts_out, soft_out, inter); break; } \
        }                                                                                                                       \
        dice_counts_kernel<TZ, TL, V, NT><<<grid, kEvThreads, 0, st>>>(p, thr, counters, partials, counts_out, soft_out, inter); \
    } while (0)
    if (ld == ECO_U8) {
        if (vec == 4) { if (zd == ECO_F32) ECO_EV(float, uint8_t, 4); else ECO_EV(__nv_bfloat16, uint8_t, 4); }
        else { if (zd == ECO_F32) ECO_EV(float, uint8_t, 1); else ECO_EV(__nv_bfloat16, uint8_t, 1); }
    } else if (vec == 4) {
        if (zd == ECO_F32 && ld == ECO_F32) ECO_EV(float, float, 4);
        else if (zd == ECO_BF16 && ld == ECO_F32) ECO_EV(__nv_bfloat16, float, 4);
        else if (zd == ECO_F32 && ld == ECO_BF16) ECO_EV(float, __nv_bfloat16, 4);
        else ECO_EV(__nv_bfloat16, __nv_bfloat16, 4);
    } else {
        if (zd == ECO_F32 && ld == ECO_F32) ECO_EV(float, float, 1);
        else if (zd == ECO_BF16 && ld == ECO_F32) ECO_EV(__nv_bfloat16, float, 1);
        else if (zd == ECO_F32 && ld == ECO_BF16) ECO_EV(float, __nv_bfloat16, 1);
        else ECO_EV(__nv_bfloat16, __nv_bfloat16, 1);
    }
#undef ECO_EV
}

static void launch_beam(const EvalArgs& p, int zd, int ld, int vec, dim3 grid, cudaStream_t st, const float* thr,
                        unsigned int* counters, long long* partials, long long* counts_out, double* soft_out, double* inter) {
#define ECO_BM(TZ, TL, V) dice_beam_kernel<TZ, TL, V><<<grid, kBeamThreads, 0, st>>>(p, thr, counters, partials, counts_out, soft_out, inter)
    if (ld == ECO_U8) {
        if (vec == 4) { if (zd == ECO_F32) ECO_BM(float, uint8_t, 4); else ECO_BM(__nv_bfloat16, uint8_t, 4); }
        else { if (zd == ECO_F32) ECO_BM(float, uint8_t, 1); else ECO_BM(__nv_bfloat16, uint8_t, 1); }
    } else if (vec == 4) {
        if (zd == ECO_F32 && ld == ECO_F32) ECO_BM(float, float, 4);
        else if (zd == ECO_BF16 && ld == ECO_F32) ECO_BM(__nv_bfloat16, float, 4);
        else if (zd == ECO_F32 && ld == ECO_BF16) ECO_BM(float, __nv_bfloat16, 4);
        else ECO_BM(__nv_bfloat16, __nv_bfloat16, 4);
    } else {
        if (zd == ECO_F32 && ld == ECO_F32) ECO_BM(float, float, 1);
        else if (zd == ECO_BF16 && ld == ECO_F32) ECO_BM(__nv_bfloat16, float, 1);
        else if (zd == ECO_F32 && ld == ECO_BF16) ECO_BM(float, __nv_bfloat16, 1);
        else ECO_BM(__nv_bfloat16, __nv_bfloat16, 1);
    }
#undef ECO_BM
}

static void launch_exact(const EvalArgs& p, int zd, int ld, int sms, cudaStream_t st, const float* thr, const unsigned int* flags,
                         double* inter) {
    dim3 grid((unsigned)(sms * 4 / p.C > 0 ? sms * 4 / p.C : 1), (unsigned)p.C, 1);
#define ECO_EX(TZ, TL) dice_exact_kernel<TZ, TL><<<grid, 256, 0, st>>>(p, thr, flags, inter)
    if (zd == ECO_F32) {
        if (ld == ECO_F32) ECO_EX(float, float); else if (ld == ECO_BF16) ECO_EX(float, __nv_bfloat16); else ECO_EX(float, uint8_t);
    } else {
        if (ld == ECO_F32) ECO_EX(__nv_bfloat16, float); else if (ld == ECO_BF16) ECO_EX(__nv_bfloat16, __nv_bfloat16); else ECO_EX(__nv_bfloat16, uint8_t);
    }
#undef ECO_EX
}

// 0 / 1 threshold: the compare kernel; 2..20: the binning kernel (round 1 ran 2..4 thresholds on the compare kernel: 297 us at
// cfg3 = 71 % of the HBM peak against 222 us = 95 % here)
static int nt_bucket(int n_thr) { return n_thr == 0 ? 0 : n_thr == 1 ? 1 : 20; }

}  // namespace eco

using namespace eco;

extern "C" int64_t eco_dice_ws_bytes(int32_t C, int32_t n_thr) {
    if (C <= 0 || n_thr < 0 || n_thr > 20) return -1;
    const int nta = nt_bucket(n_thr) > 0 ? nt_bucket(n_thr) : 1;
    return eval_ws_off(C) + (int64_t)C * kEvMaxCtas * eval_rec_words(nta) * 8;
}

extern "C" int eco_dice_counts_ex(const EcoView* logits, const EcoView* labels, int32_t N, int32_t C, int64_t HW,
                                  const float* thresholds, int32_t n_thr, int32_t logits_are_probs, void* ws,
                                  int64_t ws_bytes, int64_t* counts_out, double* soft_out, double* thr_inter_out, int device,
                                  void* stream) {
    if (N <= 0 || C <= 0 || HW <= 0) { set_error("empty input (N=%d C=%d HW=%lld)", N, C, (long long)HW); return -2; }
    if (!logits || !labels || !logits->ptr || !labels->ptr) { set_error("null input view"); return -1; }
    if (C > 65535) { set_error("C too large"); return -3; }
    if (n_thr < 0 || n_thr > 20) { set_error("n_thr must be in [0,20] per call (got %d)", n_thr); return -4; }
    if (n_thr > 0 && !thresholds) { set_error("null thresholds"); return -4; }
    if ((logits->dtype != ECO_F32 && logits->dtype != ECO_BF16) ||
        (labels->dtype != ECO_F32 && labels->dtype != ECO_BF16 && labels->dtype != ECO_U8)) {
        set_error("dice_counts: logits must be f32/bf16, labels f32/bf16/u8");
        return -4;
    }
    if (logits_are_probs & ~(ECO_EVAL_PROBS | ECO_EVAL_UNUNION)) { set_error("dice_counts: unknown flag bits 0x%x", logits_are_probs); return -4; }
    if ((logits_are_probs & ECO_EVAL_UNUNION) && n_thr > 0) {
        set_error("dice_counts: the prediction un-union is fused for the soft Dice only (n_thr = 0); un-union in place first (eco_union_sets)");
        return -4;
    }
    if (!ws || ws_bytes < eco_dice_ws_bytes(C, n_thr) || !counts_out || !soft_out) { set_error("workspace too small or null output"); return -5; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    EvalArgs p{};
    p.z = logits->ptr; p.l = labels->ptr;
    p.z_sn = logits->sn; p.z_sc = logits->sc; p.l_sn = labels->sn; p.l_sc = labels->sc;
    p.N = N; p.C = C; p.HW = HW; p.n_thr = n_thr; p.probs = (logits_are_probs & ECO_EVAL_PROBS) ? 1 : 0;
    p.unun = ((logits_are_probs & ECO_EVAL_UNUNION) && C >= 3) ? 1 : 0;
    const int vec = (ev_aligned(logits, HW) && ev_aligned(labels, HW)) ? 4 : 1;
    const int tile = kEvThreads * vec * kEvUnroll;
    p.tiles_per_plane = (int32_t)((HW + tile - 1) / tile);
    p.tiles_per_channel = (int64_t)p.tiles_per_plane * N;
    const int sms = sm_count_cached(device);
    if (sms <= 0) return -10;
    const int ctas_per_sm = nt_bucket(n_thr) == 20 ? kBeamCtasPerSm : kEvCtasPerSm;
    int64_t max_ctas = (int64_t)sms * ctas_per_sm / C;
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > kEvMaxCtas) max_ctas = kEvMaxCtas;
    int64_t per = (p.tiles_per_channel + max_ctas - 1) / max_ctas;
    p.tiles_per_cta = (int32_t)(per < 1 ? 1 : per);
    int64_t ctas = (p.tiles_per_channel + p.tiles_per_cta - 1) / p.tiles_per_cta;
    dim3 grid((unsigned)(ctas < 1 ? 1 : ctas), (unsigned)C, 1);
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    long long* partials = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + eval_ws_off(C));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long* co = reinterpret_cast<long long*>(counts_out);
    switch (nt_bucket(n_thr)) {
        case 0: launch_nt<0>(p, logits->dtype, labels->dtype, vec, grid, st, thresholds, counters, partials, co, soft_out, thr_inter_out); break;
        case 1: launch_nt<1>(p, logits->dtype, labels->dtype, vec, grid, st, thresholds, counters, partials, co, soft_out, thr_inter_out); break;
        default: launch_beam(p, logits->dtype, labels->dtype, vec, grid, st, thresholds, counters, partials, co, soft_out, thr_inter_out); break;
    }
    int rc = check_cuda(cudaGetLastError(), "dice_counts_kernel launch");
    if (rc || n_thr == 0 || !thr_inter_out) return rc;
    // exact sums for classes whose labels are not all 0/1: returns at once unless the counting kernel raised a flag
    launch_exact(p, logits->dtype, labels->dtype, sms, st, thresholds, counters + C, thr_inter_out);
    return check_cuda(cudaGetLastError(), "dice_exact_kernel launch");
}

extern "C" int eco_dice_counts(const EcoView* logits, const EcoView* labels, int32_t N, int32_t C, int64_t HW,
                               const float* thresholds, int32_t n_thr, int32_t logits_are_probs, void* ws,
                               int64_t ws_bytes, int64_t* counts_out, double* soft_out, int device, void* stream) {
    return eco_dice_counts_ex(logits, labels, N, C, HW, thresholds, n_thr, logits_are_probs, ws, ws_bytes, counts_out, soft_out,
                              nullptr, device, stream);
}

extern "C" int eco_dice_finalize_ex(const int64_t* counts, const double* soft, const double* thr_inter, int32_t C, int32_t n_thr,
                                    float* dice_out, float* soft_dice_out, int device, void* stream) {
    if (C <= 0 || n_thr < 0) { set_error("bad C/n_thr"); return -1; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    DiceFinArgs a{C, n_thr};
    const int total = (n_thr > 0 ? n_thr : 1) * C;
    dice_finalize_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(counts), soft, thr_inter, a, dice_out, soft_dice_out);
    return check_cuda(cudaGetLastError(), "dice_finalize_kernel launch");
}

extern "C" int eco_dice_finalize(const int64_t* counts, const double* soft, int32_t C, int32_t n_thr, float* dice_out,
                                 float* soft_dice_out, int device, void* stream) {
    return eco_dice_finalize_ex(counts, soft, nullptr, C, n_thr, dice_out, soft_dice_out, device, stream);
}
