// Third-generation fused 3-organ composite step (fp32 or bf16 logits; fp32 or uint8 labels; 16-byte aligned planes).
// Included by eco_composite.cu after eco_composite_v2.cuh, whose tile pipeline, pass-1 / pass-2 math and integer grid sums
// it reuses.  What changed against composite3_fused_v2_kernel, and why (measurements: profiles/README.md):
//
//   * v2 was issue-bound in pass 2 (42 us, 95 % of its issue model) and sweep-bound in pass 1 (22 us with ~40 % of the
//     issue slots idle).  About 300 of pass 2's 840 issue cycles per pixel pair were the BCE / focal LINEAR sums
//     (sum of the softplus remainder and of (1-b)^1.5 lg2(b+eps) over the 12 real-b leaves): they feed only loss VALUES,
//     never a gradient coefficient.  Here eight dedicated "linear" warps (two per SM sub-partition, 48 registers) take
//     those sums from their OWN small ring of logit tiles (a third, independent sweep over the logits: L2 hits), at their
//     own pace, from the first cycle of the kernel to wherever they finish -- under pass 1, under the grid-wide
//     hand-over (whose idle issue slots they fill) and under pass 2, which is now the bare gradient.  The 55 statistics
//     accumulators stay where they were, so nothing spills.  (First attempt, measured: four linear warps reading the
//     SAME stages as the statistics warps -- one warp per sub-partition runs at ~20 % of the issue rate, and the shared
//     ring's back-pressure stretched pass 1 from 20 to 51 us.)
//   * The linear sums are off the critical path altogether: they have their own integer accumulators and arrival
//     counter, the hand-over of the 55 statistics does not wait for them, and CTA 0 adds them to the BCE / focal totals
//     at the very end.
//   * No reduction at the kernel's tail any more (v2: one word per sum carrying value + arrival count, then -- sharded --
//     a SECOND NVLink exchange that every rank's last CTA waited for).  Sharded, the linear sums still cross NVLink, but
//     the send happens right after pass 1 and the (already satisfied) receive at the end: one exchange on the critical
//     path instead of two.
//   * The hand-over ships the 55 raw sums (+ 15 label corrections only when a label is not 0/1) instead of the derived
//     100-slot layout: 110 instead of 200 integer atomics per CTA; the layout is derived after the hand-over.
//   * Labels may be uint8 masks (ECO_U8; SURVEY 8(f)-4: the datasets produce {0,1} masks): 9 instead of 12 B/element of
//     HBM traffic and 15 instead of 24 KB per stage; and the label union of ess/utils/subsets_union.py:8-32
//     (g1 <- min(1, g1 + g2), exclude_indices=[0]) can be applied in registers at load (flag), which removes the separate
//     in-place sweep of ess/train_multiclass.py:110.
//   * Peer-exchange waits time out on %globaltimer (configurable, default 30 s) instead of a spin count, and the time-out
//     is reported through a status word the host checks (eco_xch_poll_status).
#pragma once

namespace eco {
namespace v2 {

#ifdef ECO_V2_TIMELINE
#define ECO_TLL(slot) do { if (threadIdx.x == kLinWarp0 * 32) g_timeline[blockIdx.x * 16 + (slot)] = gtime(); } while (0)
#else
#define ECO_TLL(slot) do { } while (0)
#endif

// Warp roles.  Registers are handed out per warpgroup (4 warps): the CTA is 7 warpgroups launched at 72 registers per
// thread; the four statistics / gradient warpgroups then raise their budget to 96, the two linear warpgroups lower theirs to
// 48 and the producers' warpgroup to 24 (setmaxnreg): 16 x 96 + 8 x 48 + 4 x 24 = 2016 = 28 x 72 -- what is released
// equals what is claimed, the 55 accumulators of pass 1 stay in registers.
#ifndef ECO_V3_LINWARPS
#define ECO_V3_LINWARPS 8
#endif
constexpr int kLinWarps = ECO_V3_LINWARPS;                    // two per SM sub-partition
constexpr int kLinQ = 16 / kLinWarps;                         // pixel pairs per thread and linear tile
constexpr int kLinTP = kLinWarps * 64 * kLinQ;                // pixels per tile of the linear ring (its own partition of the planes)
static_assert(kLinQ >= 1 && kLinTP == 1024, "linear tile geometry");
constexpr int kLinWarp0 = kCWarps;                            // warps 16..23
constexpr int kProdWarp = kCWarps + kLinWarps;                // warp 24: main ring; warp 25: linear ring; 26, 27 exit at once
constexpr int kThreads3 = (kCWarps + kLinWarps + 4) * 32;     // 896
__device__ __forceinline__ void reg_raise96() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory"); }
__device__ __forceinline__ void reg_lower48() { asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory"); }
__device__ __forceinline__ void reg_lower24() { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory"); }
constexpr int kStages3 = 5;
constexpr int kLinStages = 4;                                 // ring of the linear warps: logit planes only
template <typename TX>
struct LinStage {
    static constexpr int kPlane = kLinTP * (int)sizeof(TX);
    static constexpr int kBytes = 3 * kPlane;                 // 12 KB (fp32 logits) / 6 KB (bf16)
};
constexpr int kLinFlushTiles = 32 / kLinQ;                    // 64 values per fp32 partial between folds into fp64
constexpr int kFlushTiles3 = 32;                              // 64 pixels per fp32 accumulator between folds into fp64 (cfg2: one fold per CTA)
constexpr int kNFlat = 72;                                    // 55 flat sums | 15 label corrections | n | (pad)
constexpr int F_CORR = 55, F_N = 70;

constexpr unsigned int kC3FlagUnionLabels = 1u;               // == ECO_C3_UNION_LABELS
constexpr unsigned int kC3FlagNoGrad = 4u;                    // == ECO_C3_NO_GRAD: loss values only (validation), no pass 2

template <typename TX, typename TG>
struct Stage3 {
    static constexpr int kXPlane = kTP * (int)sizeof(TX);
    static constexpr int kXBytes = 3 * kXPlane;
    static constexpr int kGPlane = kTP * (int)sizeof(TG);
    static constexpr int kBytes = kXBytes + 3 * kGPlane;      // 24 KB (fp32 logits + fp32 labels) ... 9 KB (bf16 + byte labels)
    static constexpr int kMain = kStages3 * kBytes;
    static constexpr int kSmem = kMain + kLinStages * LinStage<TX>::kBytes;
};

// this thread's pixel pair of a logit plane: fp32 as is, bf16 widened exactly (a bf16 is the top half of an fp32)
template <typename TX>
__device__ __forceinline__ f2 lds_x2(uint32_t plane_addr, int pix);
template <>
__device__ __forceinline__ f2 lds_x2<float>(uint32_t plane_addr, int pix) { return lds_f2(plane_addr + (uint32_t)pix * 4); }
template <>
__device__ __forceinline__ f2 lds_x2<__nv_bfloat16>(uint32_t plane_addr, int pix) {
    unsigned int w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(plane_addr + (uint32_t)pix * 2));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// gradient of a pixel pair out: streaming store, bf16 rounded to nearest even
__device__ __forceinline__ void stg_grad2(float* p, f2 v) { stg_stream_f2(p, v); }
__device__ __forceinline__ void stg_grad2(__nv_bfloat16* p, f2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(*reinterpret_cast<const unsigned int*>(&h)) : "memory");
}

struct PipeSmem3 {
    unsigned long long full[kStages3];
    unsigned long long empty[kStages3];
    unsigned long long lfull[kLinStages];
    unsigned long long lempty[kLinStages];
};

__device__ __forceinline__ void pipe_init3(PipeSmem3& ps) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages3; ++s) {
            mbar_init(smem_u32(&ps.full[s]), 1);
            mbar_init(smem_u32(&ps.empty[s]), kCWarps);
        }
#pragma unroll
        for (int s = 0; s < kLinStages; ++s) {
            mbar_init(smem_u32(&ps.lfull[s]), 1);
            mbar_init(smem_u32(&ps.lempty[s]), kLinWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

template <typename TX, typename TG>
__device__ __forceinline__ void produce_tiles3(const CompArgs& a, const TileRange& tr, bool reverse, uint32_t stage_base,
                                               PipeSmem3& ps, int k0) {
    const TX* xb = reinterpret_cast<const TX*>(a.x);
    const TG* gb = reinterpret_cast<const TG*>(a.g);
    const int ntiles = tr.t_hi - tr.t_lo;
    if (ntiles <= 0) return;
    int t = reverse ? tr.t_hi - 1 : tr.t_lo;
    int n = t / tr.tpp, kk = t - n * tr.tpp;
    for (int k = 0; k < ntiles; ++k) {
        const int kg = k0 + k;
        const int s = kg % kStages3;
        const uint32_t full = smem_u32(&ps.full[s]), empty = smem_u32(&ps.empty[s]);
        if (kg >= kStages3) mbar_wait(empty, ((kg / kStages3) - 1) & 1);
        const int64_t p0 = (int64_t)kk * kTP;
        const int valid = (int)((a.HW - p0 < kTP) ? (a.HW - p0) : kTP);
        const uint32_t xbytes = (uint32_t)valid * (uint32_t)sizeof(TX), gbytes = (uint32_t)valid * (uint32_t)sizeof(TG);
        const uint32_t dst = stage_base + (uint32_t)s * Stage3<TX, TG>::kBytes;
        const TX* xs = xb + n * a.x_sn + p0;
        const TG* gs = gb + n * a.g_sn + p0;
        mbar_expect_tx(full, 3u * (xbytes + gbytes));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bulk_g2s(dst + (uint32_t)c * Stage3<TX, TG>::kXPlane, xs + c * a.x_sc, xbytes, full);
            bulk_g2s(dst + Stage3<TX, TG>::kXBytes + (uint32_t)c * Stage3<TX, TG>::kGPlane, gs + c * a.g_sc, gbytes, full);
        }
        if (reverse) { if (--kk < 0) { kk = tr.tpp - 1; --n; } }
        else { if (++kk == tr.tpp) { kk = 0; ++n; } }
    }
}

// this thread's pixel pair of label plane c of a stage
template <typename TG>
__device__ __forceinline__ f2 lds_label2(uint32_t labels_addr, int c);   // labels_addr = first label plane of the stage
template <>
__device__ __forceinline__ f2 lds_label2<float>(uint32_t labels_addr, int c) {
    return lds_f2(labels_addr + (uint32_t)c * (kTP * 4) + threadIdx.x * 8);
}
template <>
__device__ __forceinline__ f2 lds_label2<uint8_t>(uint32_t labels_addr, int c) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(labels_addr + (uint32_t)c * kTP + threadIdx.x * 2));
    // byte -> float without I2F (which shares the XU pipe with the sigmoids): 0x4b000000 | b is 8388608 + b exactly
    const unsigned int w = v;
    return make_float2(__uint_as_float(0x4b000000u | (w & 0xffu)) - 8388608.0f, __uint_as_float(0x4b000000u | (w >> 8)) - 8388608.0f);
}

// label union of ess/utils/subsets_union.py:8-32 with exclude_indices=[0] on 3 classes: channel 1 <- sum of channels
// 1.. (:26), then EVERY entry above 1 is set to 1 (:28)
__device__ __forceinline__ float clamp_above1(float v) { return v > 1.0f ? 1.0f : v; }
__device__ __forceinline__ void union_labels(f2 (&g)[3]) {
    g[1] = add2(g[1], g[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = make_float2(clamp_above1(g[c].x), clamp_above1(g[c].y));
}

template <typename TX, typename TG>
__device__ __forceinline__ void consume_tile3(uint32_t stage_base, PipeSmem3& ps, int kg, int lane, bool uni, f2 (&z)[3], f2 (&g)[3]) {
    const int s = kg % kStages3;
    mbar_wait(smem_u32(&ps.full[s]), (kg / kStages3) & 1);
    const uint32_t sb = stage_base + (uint32_t)s * Stage3<TX, TG>::kBytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        z[c] = lds_x2<TX>(sb + (uint32_t)c * Stage3<TX, TG>::kXPlane, 2 * (int)threadIdx.x);
        g[c] = lds_label2<TG>(sb + Stage3<TX, TG>::kXBytes, c);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&ps.empty[s]));
    if (uni) union_labels(g);
}

// ---------------------------------------------------------------------------------------------
// peer exchange (NCCL-LL style, see ll_send / ll_recv_sum of v2) with an explicit slot per call and a wall-clock time-out
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ inline void ll_send1(const XchArgs& x, double mine, int slot) {
    const int par = x.epoch & 1u;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
    const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
    for (int r = 0; r < x.world; ++r) {
        char* dst = reinterpret_cast<char*>(x.peers[r]) + xch_ll_offset_bytes(x.world) +
                    ((size_t)(par * x.world + x.rank)) * (128 * 16) + (size_t)slot * 16;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(lo), "r"(x.epoch), "r"(hi), "r"(x.epoch) : "memory");
    }
}
__device__ inline double ll_recv1(const XchArgs& x, int slot) {
    const int par = x.epoch & 1u;
    double tot = 0.0;
    const char* own = reinterpret_cast<const char*>(x.peers[x.rank]) + xch_ll_offset_bytes(x.world) +
                      ((size_t)(par * x.world)) * (128 * 16) + (size_t)slot * 16;
    unsigned long long t0 = 0ull;
    for (int r = 0; r < x.world; ++r) {
        unsigned int lo, f0, hi, f1, spins = 0;
        while (true) {
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(own + (size_t)r * (128 * 16)) : "memory");
            if (f0 == x.epoch && f1 == x.epoch) break;
            if ((++spins & 1023u) == 0u) {   // look at the clock every ~1000 polls
                const unsigned long long now = gtimer_ns();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > x.timeout_ns) {
                    *x.status = 1u;                  // a peer never showed up: poison (NaN) instead of hanging the GPU;
                    lo = 0u; hi = 0x7ff80000u;       // the host sees the status word (eco_xch_poll_status) and raises
                    break;
                }
            }
            __nanosleep(20);
        }
        tot += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
    return tot;
}

// ---------------------------------------------------------------------------------------------
// workspace of the v3 kernel (all zero before the first launch).  Everything is double buffered by step parity: step k
// works in buffer k & 1 while CTA 0 clears buffer (k + 1) & 1 for the next step, so there is no re-arming phase at the
// end of a step and nobody ever waits for a clear.
// ---------------------------------------------------------------------------------------------
struct V3Ws {
    unsigned int step, _pad[3];
    unsigned int arrive1[2][4];        // [par][0]: CTAs whose statistics are in fix1
    unsigned int arrive_lin[2][4];     // [par][0]: CTAs whose linear sums are in lin
    unsigned long long lin[2][kFixRep][4];           // 2 sums x (hi, lo)
    unsigned long long fix1[2][kFixRep][2 * kNFlat];
};

struct Fused3Smem {
    StatsSmem st;
    LeafCoef cf[ECO_C3_NLEAF];
    Coef2 c2;
    double sl[ECO_C3_NLEAF][ECO_NLOSS];
    double jac_s[ECO_C3_NLEAF][ECO_NLOSS][ECO_NJAC];
    double scale[ECO_C3_NLEAF];     // linear warps' copy
    double scale_c[ECO_C3_NLEAF];   // statistics warps' copy
    double flat[kNFlat];
    double lin_part[kLinWarps][2];
    double lin_tot[2];
    float up[ECO_NLOSS + 1];
    PipeSmem3 ps;
};

// The linear sums carry the caller's leaf scales (up to ~23.3 |w| per pixel and leaf), so their integer accumulators use a
// coarser high word than fix_add: |sum| < 2^53, still exact to 2^-42.
constexpr double kLinHi = 1024.0;   // 2^10
__device__ __forceinline__ void lin_fix_add(unsigned long long* slot2, double v) {
    const double sc = v * kLinHi;
    const long long hi = __double2ll_rn(sc);
    const long long lo = __double2ll_rn((sc - (double)hi) * kFixLo);
    atomicAdd(slot2, (unsigned long long)hi);
    atomicAdd(slot2 + 1, (unsigned long long)lo);
}
__device__ __forceinline__ double lin_fix_get(const unsigned long long* slot2, int rep_stride) {
    return fix_get(slot2, rep_stride) * (kFixHi / kLinHi);
}

// barrier over the linear warps only
__device__ __forceinline__ void lsync() { asm volatile("bar.sync 2, %0;" ::"n"(kLinWarps * 32) : "memory"); }

// producer of the linear ring: the logit planes of this CTA's tiles, forwards, once
template <typename TX>
__device__ __forceinline__ void produce_lin_tiles(const CompArgs& a, const TileRange& tr, uint32_t lin_base, PipeSmem3& ps) {
    const TX* xb = reinterpret_cast<const TX*>(a.x);
    const int ntiles = tr.t_hi - tr.t_lo;
    if (ntiles <= 0) return;
    int n = tr.t_lo / tr.tpp, kk = tr.t_lo - n * tr.tpp;
    for (int k = 0; k < ntiles; ++k) {
        const int s = k % kLinStages;
        const uint32_t full = smem_u32(&ps.lfull[s]);
        if (k >= kLinStages) mbar_wait(smem_u32(&ps.lempty[s]), ((k / kLinStages) - 1) & 1);
        const int64_t p0 = (int64_t)kk * kLinTP;
        const int valid = (int)((a.HW - p0 < kLinTP) ? (a.HW - p0) : kLinTP);
        const uint32_t xbytes = (uint32_t)valid * (uint32_t)sizeof(TX);
        const uint32_t dst = lin_base + (uint32_t)s * LinStage<TX>::kBytes;
        const TX* xs = xb + n * a.x_sn + p0;
        mbar_expect_tx(full, 3u * xbytes);
#pragma unroll
        for (int c = 0; c < 3; ++c) bulk_g2s(dst + (uint32_t)c * LinStage<TX>::kPlane, xs + c * a.x_sc, xbytes, full);
        if (++kk == tr.tpp) { kk = 0; ++n; }
    }
}

// the eight linear warps: BCE / focal linear sums of every tile of this CTA, from their own ring
template <typename TX, bool POSW, bool PROB>
__device__ __forceinline__ void lin_consume(const CompArgs& a, const TileRange& tr, uint32_t lin_base, PipeSmem3& ps,
                                            const Coef2& c2, double (&tot)[2]) {
    const int lane = threadIdx.x & 31, lw = (threadIdx.x >> 5) - kLinWarp0;
    const int ntiles = tr.t_hi - tr.t_lo;
    int kk = ntiles > 0 ? tr.t_lo % tr.tpp : 0;
    f2 sp_acc = splat(0.f), fl_acc = splat(0.f);
    int since = 0;
    tot[0] = tot[1] = 0.0;
    for (int k = 0; k < ntiles; ++k) {
        const int s = k % kLinStages;
        mbar_wait(smem_u32(&ps.lfull[s]), (k / kLinStages) & 1);
        const uint32_t sb = lin_base + (uint32_t)s * LinStage<TX>::kBytes;
        const int64_t p0 = (int64_t)kk * kLinTP;
#pragma unroll 1
        for (int q = 0; q < kLinQ; ++q) {
            const int pix = ((q * kLinWarps + lw) * 32 + lane) * 2;
#ifndef ECO_V3_EXP_LIN_OFF
            if (p0 + pix < a.HW) {
                f2 x[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const f2 v = lds_x2<TX>(sb + (uint32_t)c * LinStage<TX>::kPlane, pix);
                    x[c] = PROB ? v : sigmoid_fast2(v);
                }
                pixel_pair_tr<POSW>(x, c2, sp_acc, fl_acc);
            }
#else
            (void)pix; (void)p0; (void)sb;
#endif
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ps.lempty[s]));
        if (++kk == tr.tpp) kk = 0;
        if (++since == kLinFlushTiles) {
            tot[0] += (double)(sp_acc.x + sp_acc.y);
            tot[1] += (double)(fl_acc.x + fl_acc.y);
            sp_acc = splat(0.f); fl_acc = splat(0.f);
            since = 0;
        }
    }
    tot[0] += (double)(sp_acc.x + sp_acc.y);
    tot[1] += (double)(fl_acc.x + fl_acc.y);
}

// pass 1, statistics warps: as stats_consume of v2 on the v3 stage layout
template <typename TX, typename TG, bool PROB>
__device__ __forceinline__ void stats_consume3(const CompArgs& a, const TileRange& tr, uint32_t stage_base, PipeSmem3& ps,
                                               bool uni, StatsSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntiles = tr.t_hi - tr.t_lo;
    int kk = ntiles > 0 ? tr.t_lo % tr.tpp : 0;
    float acc[F_NACC];
#pragma unroll
    for (int k = 0; k < F_NACC; ++k) acc[k] = 0.f;
    int since_flush = 0;
    bool any_nonbinary = false;
    for (int k = 0; k < ntiles; ++k) {
        f2 z[3], g[3];
        consume_tile3<TX, TG>(stage_base, ps, k, lane, uni, z, g);
        if ((int64_t)kk * kTP + 2 * (int)threadIdx.x < a.HW) {
            stats_pixel<PROB>(z[0].x, z[1].x, z[2].x, g[0].x, g[1].x, g[2].x, acc);
            stats_pixel<PROB>(z[0].y, z[1].y, z[2].y, g[0].y, g[1].y, g[2].y, acc);
        }
        if (++kk == tr.tpp) kk = 0;
        if (++since_flush == kFlushTiles3) {
            any_nonbinary |= flush_flat_acc(acc, sm.warp_slots[warp], lane);
            since_flush = 0;
        }
    }
    any_nonbinary |= flush_flat_acc(acc, sm.warp_slots[warp], lane);
    if (any_nonbinary) sm.flag = true;
}

// CTA-level tail of pass 1 (statistics warps): rare slow pass, then the CTA's raw sums go into the integer accumulators
// and the CTA arrives.  Returns true in the last CTA to arrive.
template <typename TG>
__device__ __forceinline__ bool stats_finish3(const CompArgs& a, const TileRange& tr, bool uni, StatsSmem& sm, V3Ws* ws, int par) {
    csync();
    const bool nonbinary = sm.flag;
    if (nonbinary) {
        const TG* gb = reinterpret_cast<const TG*>(a.g);
        for (int t = tr.t_lo; t < tr.t_hi; ++t) {
            const int n = t / tr.tpp, kk = t - n * tr.tpp;
            const int64_t p0 = (int64_t)kk * kTP;
            for (int e = threadIdx.x; e < kTP && p0 + e < a.HW; e += kCThreads) {
                const TG* gp = gb + n * a.g_sn + p0 + e;
                float g0 = (float)gp[0], g1 = (float)gp[a.g_sc], g2 = (float)gp[2 * a.g_sc];
                if (uni) { g1 = clamp_above1(g1 + g2); g0 = clamp_above1(g0); g2 = clamp_above1(g2); }
                if ((g0 != 0.f && g0 != 1.f) || (g1 != 0.f && g1 != 1.f) || (g2 != 0.f && g2 != 1.f))
                    label_corrections_v2(g0, g1, g2, sm.corr);
            }
        }
        csync();
    }
    if (threadIdx.x < 64) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kCWarps; ++w) v += sm.warp_slots[w][threadIdx.x];
        sm.sums[threadIdx.x] = v;
    }
    csync();
    if (threadIdx.x < kNFlat) {
        double v = 0.0;
        bool send = false;
        if (threadIdx.x < F_NACC) { v = sm.sums[threadIdx.x]; send = true; }
        else if (threadIdx.x < F_N) { v = sm.corr[threadIdx.x - F_CORR]; send = nonbinary; }
        else if (threadIdx.x == F_N) {
            const int64_t last = a.HW - (int64_t)(tr.tpp - 1) * kTP;   // pixels of the (possibly short) last tile of a plane
            const int ntiles = tr.t_hi - tr.t_lo;
            const int n_last = ntiles > 0 ? (tr.t_hi / tr.tpp - tr.t_lo / tr.tpp) : 0;
            v = (double)((int64_t)(ntiles - n_last) * kTP + (int64_t)n_last * last);
            send = true;
        }
        if (send) {
            fix_add(ws->fix1[par][blockIdx.x % kFixRep] + 2 * threadIdx.x, v);
            __threadfence();
        }
    }
    csync();
    ECO_TL(15);
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&ws->arrive1[par][0], 1u);
        sm.flag = (prev == gridDim.x - 1);
    }
    csync();
    return sm.flag;
}

// pass 2, gradient warps: grad_consume of v2 on the v3 stage layout, walking this CTA's tiles backwards (the lines pass 1
// left in L2 come first); no linear sums here any more
template <typename TX, typename TG, bool SIG, bool FL, bool PROB>
__device__ __forceinline__ void grad_consume3(const CompGradArgs& ga, const TileRange& tr, uint32_t stage_base, PipeSmem3& ps,
                                              int k0, bool uni, const Coef2& c2, const LeafCoef* cf) {
    const CompArgs& a = ga.a;
    TX* __restrict__ ob = reinterpret_cast<TX*>(ga.gx);
    const int lane = threadIdx.x & 31;
    const int ntiles = tr.t_hi - tr.t_lo;
    if (ntiles <= 0) return;
    int t = tr.t_hi - 1;
    int n = t / tr.tpp, kk = t - n * tr.tpp;
    const int pix = 2 * (int)threadIdx.x;
    for (int k = 0; k < ntiles; ++k) {
        f2 z[3], g[3];
        consume_tile3<TX, TG>(stage_base, ps, k0 + k, lane, uni, z, g);
        const int64_t p0 = (int64_t)kk * kTP;
        if (p0 + pix < a.HW) {
            f2 x[3], gx[3], diffs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) x[c] = PROB ? z[c] : sigmoid_fast2(z[c]);
#pragma unroll
            for (int p = 0; p < 3; ++p) diffs[p] = add2(x[pair_i(p)], neg2(x[pair_j(p)]));
            pixel_pair_grad2<SIG, FL>(x, g, diffs, c2, gx);
            f2 o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = PROB ? gx[c] : mul2(gx[c], mul2(x[c], fma2(x[c], splat(-1.0f), splat(1.0f))));
            const float dx = fminf(fminf(fabsf(diffs[0].x), fabsf(diffs[1].x)), fabsf(diffs[2].x));
            const float dy = fminf(fminf(fabsf(diffs[0].y), fabsf(diffs[1].y)), fabsf(diffs[2].y));
            if (fminf(dx, dy) < kTieEps) {
                if (dx < kTieEps) {
                    const float3 r = tie_pixel_grad(z[0].x, z[1].x, z[2].x, g[0].x, g[1].x, g[2].x, cf, SIG, FL, PROB);
                    o[0].x = r.x; o[1].x = r.y; o[2].x = r.z;
                }
                if (dy < kTieEps) {
                    const float3 r = tie_pixel_grad(z[0].y, z[1].y, z[2].y, g[0].y, g[1].y, g[2].y, cf, SIG, FL, PROB);
                    o[0].y = r.x; o[1].y = r.y; o[2].y = r.z;
                }
            }
            TX* op = ob + n * ga.gx_sn + p0 + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) stg_grad2(op + c * ga.gx_sc, o[c]);
        }
        if (--kk < 0) { kk = tr.tpp - 1; --n; }
    }
}

// raw sums + corrections + n -> the 8 statistics of one of the 21 leaves, directly (composite_leaf_sums of
// eco_composite.cu applied to flat_to_layout, without materialising the 100-slot layout: one phase and one barrier less on
// the critical path between the passes).  The SP entries of the real-b leaves hold only the algebraic part
// n ln2 + sum b / 2 + sum b^2 / 8 and their FL entries are 0: the linear warps supply the rest.
__device__ inline void flat_leaf_sums(const double* F, int leaf, double* s /*[8]*/) {
    const double n = F[F_N];
    s[S_N] = n;
    s[S_FLB] = 0.0;
    if (leaf < 3) {
        s[S_A] = F[F_G + leaf]; s[S_B] = F[F_X + leaf]; s[S_BB] = F[F_XX + leaf]; s[S_AB] = F[F_GX + leaf];
        s[S_SP] = n * kLn2d + 0.5 * s[S_B] + 0.125 * s[S_BB];
        s[S_FL] = 0.0;
        return;
    }
    const int p = (leaf - 3) / 6, t = (leaf - 3) % 6, grp = t >> 1;
    const int i = pair_i(p), j = pair_j(p);
    const double* r = F + F_PAIR + 14 * p;
    const double m = r[2 + 2 * grp];
    if (t & 1) {   // U-leaf: a = g_i, b = u_k = x_i + (p_k - x_i p_k) / 2
        const double psum = grp == 0 ? F[F_X + j] : (grp == 1 ? r[1] : r[4]);
        s[S_A] = F[F_G + i];
        s[S_B] = F[F_X + i] + 0.5 * (psum - m);
        s[S_BB] = r[8 + 2 * grp];
        s[S_AB] = r[9 + 2 * grp];
        s[S_SP] = n * kLn2d + 0.5 * s[S_B] + 0.125 * s[S_BB];
        s[S_FL] = 0.0;
    } else {       // I-leaf: a = m_k, b = label (g_j for I1, |g_i - g_j| for I2 / I3)
        const double sb = (t == 0) ? F[F_G + j] : r[0];
        const double* co = F + F_CORR + 3 * ((t == 0) ? (j - 1) : (2 + p));
        s[S_A] = m; s[S_AB] = r[3 + 2 * grp]; s[S_B] = sb;
        s[S_BB] = sb + co[0];
        s[S_SP] = (n - sb) * kSP0 + sb * kSP1 + co[1];
        s[S_FL] = (n - sb) * kFL0 + co[2];
    }
}

template <typename TX, typename TG, bool PROB>
__global__ void __launch_bounds__(kThreads3, 1)
composite3_fused_v3_kernel(CompGradArgs ga, const double* __restrict__ scale_dev, const float* __restrict__ upstream,
                           V3Ws* __restrict__ ws, float* __restrict__ losses_out, unsigned int flags, XchArgs xch,
                           const float* __restrict__ upstream_prev) {
    extern __shared__ __align__(128) char stage_smem[];
    __shared__ Fused3Smem fs;
    const int warp = threadIdx.x >> 5;
    if (upstream_prev) {
        // "only if changed": the outputs already hold the step for `upstream_prev`; every thread of every CTA compares the
        // two weight vectors (same memory, same answer) and the whole grid leaves before it touches anything
        bool same = true;
#pragma unroll
        for (int k = 0; k < ECO_NLOSS; ++k) same = same && (__float_as_uint(upstream[k]) == __float_as_uint(upstream_prev[k]));
        if (same) return;
    }
    stats_smem_init(fs.st);
    pipe_init3(fs.ps);
    const TileRange tr = tile_range(ga.a);
    const TileRange ltr = tile_range(ga.a, kLinTP);   // the linear warps partition the planes on their own
    const int ntiles = tr.t_hi - tr.t_lo;
    const uint32_t sbase = smem_u32(stage_smem);
    const bool uni = (flags & kC3FlagUnionLabels) != 0u;
    const bool no_grad = (flags & kC3FlagNoGrad) != 0u;
    if (warp >= kProdWarp) {
        // ---- producers (no path from here joins another role's code: the register budgets differ)
        reg_lower24();
        if (threadIdx.x == kProdWarp * 32) {
            // main ring: pass 1 forwards, then straight on to pass 2 backwards
            produce_tiles3<TX, TG>(ga.a, tr, false, sbase, fs.ps, 0);
            if (!no_grad) produce_tiles3<TX, TG>(ga.a, tr, true, sbase, fs.ps, ntiles);
        } else if (threadIdx.x == (kProdWarp + 1) * 32) {
            produce_lin_tiles<TX>(ga.a, ltr, sbase + Stage3<TX, TG>::kMain, fs.ps);
        }
        return;
    }
    if (warp >= kLinWarp0) {
        // ---- linear warps ---------------------------------------------------------------------------------------
        reg_lower48();
        const int par = (int)(__ldcg(&ws->step) & 1u);   // advanced by CTA 0 at the end of the previous step
        const int lt = threadIdx.x - kLinWarp0 * 32;
        if (lt < ECO_C3_NLEAF) fs.scale[lt] = scale_dev[lt];
        lsync();
        bool posw = true;   // every real-b leaf has a non-negative scale: the focal weight rides on 1 - b
        for (int t = 0; t < ECO_C3_NLEAF; ++t) posw = posw && (u_leaf_of(t) < 0 || fs.scale[t] >= 0.0);
        fill_weights(fs.c2, fs.scale, lt, posw);
        lsync();
        double tot[2];
        ECO_TLL(8);
        if (posw) lin_consume<TX, true, PROB>(ga.a, ltr, sbase + Stage3<TX, TG>::kMain, fs.ps, fs.c2, tot);
        else lin_consume<TX, false, PROB>(ga.a, ltr, sbase + Stage3<TX, TG>::kMain, fs.ps, fs.c2, tot);
        ECO_TLL(9);
        tot[0] = warp_sum(tot[0]);
        tot[1] = warp_sum(tot[1]);
        if ((threadIdx.x & 31) == 0) { fs.lin_part[warp - kLinWarp0][0] = tot[0]; fs.lin_part[warp - kLinWarp0][1] = tot[1]; }
        lsync();
        if (lt < 2) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kLinWarps; ++w) v += fs.lin_part[w][lt];
            lin_fix_add(ws->lin[par][blockIdx.x % kFixRep] + 2 * lt, v);
            __threadfence();
        }
        lsync();
        if (lt == 0) {
            const unsigned int prev = atomicAdd(&ws->arrive_lin[par][0], 1u);
            if (prev == gridDim.x - 1 && xch.world > 1) {
                // this rank's linear sums are complete: send them now, CTA 0 receives at the end of the step
                __threadfence();
                ll_send1(xch, lin_fix_get(ws->lin[par][0] + 0, 4), 100);
                ll_send1(xch, lin_fix_get(ws->lin[par][0] + 2, 4), 101);
            }
        }
        ECO_TLL(10);
        return;
    }
    reg_raise96();
    ECO_TL(0);
    const int par = (int)(__ldcg(&ws->step) & 1u);   // advanced by CTA 0 at the end of the previous step
    // ---- statistics / gradient warps ---------------------------------------------------------------------------
    if (blockIdx.x == 0) {   // clear the other parity's buffers for the next step (nobody touches them during this one)
        for (int i = threadIdx.x; i < kFixRep * 2 * kNFlat; i += kCThreads) (&ws->fix1[par ^ 1][0][0])[i] = 0ull;
        if (threadIdx.x < kFixRep * 4) (&ws->lin[par ^ 1][0][0])[threadIdx.x] = 0ull;
        if (threadIdx.x == 0) { ws->arrive1[par ^ 1][0] = 0u; ws->arrive_lin[par ^ 1][0] = 0u; }
    }
    if (threadIdx.x < ECO_NLOSS) fs.up[threadIdx.x] = upstream[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + ECO_C3_NLEAF) fs.scale_c[threadIdx.x - 32] = scale_dev[threadIdx.x - 32];
    stats_consume3<TX, TG, PROB>(ga.a, tr, sbase, fs.ps, uni, fs.st);
    ECO_TL(1);
    const bool last1 = stats_finish3<TG>(ga.a, tr, uni, fs.st, ws, par);
    ECO_TL(2);
    // grid-wide hand-over of the raw sums (all CTAs are co-resident: cooperative launch)
    if (xch.world <= 1) {
        if (threadIdx.x == 0) while (ld_acquire_gpu(&ws->arrive1[par][0]) < gridDim.x) __nanosleep(32);
        csync();
        if (threadIdx.x < kNFlat) fs.flat[threadIdx.x] = fix_get(ws->fix1[par][0] + 2 * threadIdx.x, 2 * kNFlat);
    } else {
        // sharded: the last CTA to arrive sends this rank's totals to every rank (its own included); EVERY CTA then
        // receives the `world` rows itself -- one NVLink hop, no second hand-over inside the GPU
        if (last1) {
            __threadfence();
            if (threadIdx.x < kNFlat) ll_send1(xch, fix_get(ws->fix1[par][0] + 2 * threadIdx.x, 2 * kNFlat), threadIdx.x);
        }
        if (threadIdx.x < kNFlat) fs.flat[threadIdx.x] = ll_recv1(xch, threadIdx.x);
    }
    csync();
    ECO_TL(3);
    // closed forms, redundantly per CTA: one thread per (leaf, loss) row, one WARP per loss kind (no divergent switch);
    // each row leaves its Jacobian already weighted by its upstream gradient
    if (threadIdx.x < ECO_NLOSS * 32 && (threadIdx.x & 31) < ECO_C3_NLEAF) {
        const int leaf = threadIdx.x & 31, k = threadIdx.x >> 5;
        double s[ECO_NSTAT], jrow[ECO_NJAC];
        flat_leaf_sums(fs.flat, leaf, s);
        leaf_closed_form_row(s, 0.0, fs.scale_c[leaf], k, fs.sl[leaf][k], jrow);
        const float u = k == 0 ? 0.f : fs.up[k];
#pragma unroll
        for (int j = 0; j < ECO_NJAC; ++j) fs.jac_s[leaf][k][j] = u != 0.f ? (double)u * jrow[j] : 0.0;   // (an unused loss may have a non-finite Jacobian)
    }
    csync();
    ECO_TL(12);
    if (threadIdx.x < ECO_C3_NLEAF) {
        // coefficient j of leaf l = sum_k upstream[k] * d loss_k / d stat_j, in a fixed order
        const int leaf = threadIdx.x;
        float c[ECO_NJAC];
#pragma unroll
        for (int j = 0; j < ECO_NJAC; ++j) {
            double v = 0.0;
#pragma unroll
            for (int k = 1; k < ECO_NLOSS; ++k) v += fs.jac_s[leaf][k][j];
            c[j] = (float)(j == 3 ? 2.0 * v : v);
        }
        LeafCoef lc;
        lc.sa = c[0]; lc.sb = c[1]; lc.sab = c[2]; lc.sbb2 = c[3]; lc.sp = c[4]; lc.fl = c[5]; lc.flb = c[6];
        fs.cf[leaf] = lc;
        fill_coef2(fs.c2, fs.cf, leaf);   // (reads back this thread's own entry)
    }
    csync();
    ECO_TL(4);
    if (!no_grad) {
        const bool need_sig = fs.up[1] != 0.f, need_fl = fs.up[2] != 0.f;
        if (need_fl) {
            if (need_sig) grad_consume3<TX, TG, true, true, PROB>(ga, tr, sbase, fs.ps, ntiles, uni, fs.c2, fs.cf);
            else grad_consume3<TX, TG, false, true, PROB>(ga, tr, sbase, fs.ps, ntiles, uni, fs.c2, fs.cf);
        } else {
            if (need_sig) grad_consume3<TX, TG, true, false, PROB>(ga, tr, sbase, fs.ps, ntiles, uni, fs.c2, fs.cf);
            else grad_consume3<TX, TG, false, false, PROB>(ga, tr, sbase, fs.ps, ntiles, uni, fs.c2, fs.cf);
        }
    }
    ECO_TL(5);
    if (blockIdx.x != 0) return;
    // CTA 0 finishes the loss values: the linear sums arrived long ago (right after pass 1)
    if (threadIdx.x < 2) {
        double v;
        if (xch.world > 1) {
            v = ll_recv1(xch, 100 + threadIdx.x);   // every rank's last linear warp sent its totals to every rank
        } else {
            while (ld_acquire_gpu(&ws->arrive_lin[par][0]) < gridDim.x) __nanosleep(32);
            v = lin_fix_get(ws->lin[par][0] + 2 * threadIdx.x, 4);
        }
        fs.lin_tot[threadIdx.x] = v;
    }
    csync();
    if (threadIdx.x < ECO_NLOSS) {
        double v = 0.0;
        for (int l = 0; l < ECO_C3_NLEAF; ++l) v += fs.sl[l][threadIdx.x];
        const double n = fs.flat[F_N];
        if (threadIdx.x == 1) v += fs.lin_tot[0] / n;               // BCE: sum_l scale_l * softplus remainder
        if (threadIdx.x == 2) v += -kLn2d * fs.lin_tot[1] / n;      // focal: sums were taken in log2 units
        losses_out[threadIdx.x] = (float)v;
    }
    if (threadIdx.x == 0) ws->step = (unsigned int)par + 1u;   // every CTA read `step` before it arrived in pass 1
    ECO_TL(6);
}

}  // namespace v2
}  // namespace eco
