// Second-generation 3-organ composite building blocks (fp32, from-logits, 16-byte aligned planes): the TMA tile pipeline,
// the scalar pass-1 statistics, the packed pass-2 gradient, the integer grid sums, and the stand-alone gradient kernel
// composite3_grad_v2_kernel.  Included by eco_composite.cu after eco_composite_packed.cuh; the fused step built from these
// blocks is composite3_fused_v3_kernel (eco_composite_v3.cuh) -- its predecessor composite3_fused_v2_kernel, which took the
// linear BCE / focal sums in pass 2 and finished them with a second reduction at the kernel's tail, was removed in round 2
// (78 -> 72 us at cfg2, one NVLink exchange instead of two).  Ragged / unaligned planes and bf16 probabilities keep the
// first-generation kernels.
//
// Design notes (all numbers measured on B200; profiles/microbench/regbw2.cu, profiles/README.md):
//   * An SM sub-partition issues one warp instruction per clock and an instruction holds the issue path for
//     max(1, its register-file read cycles): FFMA2/FMUL2/FADD2 take 2 cycles with <= 2 distinct vector-register
//     operands (immediates, ".F32" scalar pairs and reuse-cache hits are free) and 3 cycles with 3; MUFU, ALU, LDS
//     and branches take one cycle each and do NOT overlap with FMA issue.  The kernel's time is therefore the SUM of
//     those cycles, and the loops below are written to minimise it rather than to "balance pipes":
//     universal polynomial constants are immediates, per-leaf coefficients are scalars, the chain rule is factored
//     through the intermediates d = |x_i - x_j| and q = x_i d (17 instead of 32 FMA-class instructions per pair).
//   * Pass 1 is SCALAR and not role-split: one thread owns all 55 sums of its pixels, so each sigmoid is computed
//     once (the role-split packed kernel computes it twice) and the float2 packing, which buys no issue cycles,
//     is dropped where it would cost 2x the accumulator registers.
//   * BCE and focal are LINEAR in their per-pixel sums, so those sums do not feed the gradient coefficients: they are
//     taken apart from the statistics (pixel_pair_tr: two scalars already weighted by the leaf scales, folded into the
//     polynomial coefficients and into the focal base) by the fused kernel's dedicated warps.
//   * Both passes read their input through a ring of shared-memory stages filled by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) issued by a dedicated producer warp that runs ahead from pass 1 into
//     pass 2; 16 consumer warps (4 per sub-partition, 96 registers) do nothing but math.  Pass 2 walks the CTA's tiles
//     backwards so that it starts on the lines pass 1 left in L2.
//   * Grid-wide sums are 64-bit INTEGER atomics (exact to 2^-62, order-independent => deterministic, no serial
//     "last CTA adds everything"); every CTA polls the arrival counter itself (cooperative launch: co-resident) and
//     derives the closed forms.
//   * Sharded (one process per GPU): the rank totals cross NVLink as self-validating 16-byte stores (NCCL LL style),
//     sent by the CTA that completes the rank's sums and received by every CTA of every rank.
//   * Tied pixels (|x_i - x_j| < 4e-6, ~2e-5 of all pixels) are recomputed in pass 2 by a scalar slow path with
//     ATen's exact sigmoid bits and sign(0) = 0; the sign of the |.| kink is otherwise one LOP3 per lane.
#pragma once

namespace eco {
namespace v2 {

#ifdef ECO_V2_TIMELINE
__device__ unsigned long long g_timeline[1024 * 16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ECO_TL(slot) do { if (threadIdx.x == 0) g_timeline[blockIdx.x * 16 + (slot)] = gtime(); } while (0)
#else
#define ECO_TL(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ f2 lds_f2(uint32_t a) {
    f2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void stg_stream_f2(float* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Tile pipeline.  A tile = kTP consecutive pixels of one image x 6 planes (x0 x1 x2 g0 g1 g2); it never straddles
// two images and the last tile of a plane may be short.  One consumer thread owns one pixel PAIR of a tile.
// ---------------------------------------------------------------------------------------------
#ifndef ECO_V2_CWARPS
#define ECO_V2_CWARPS 16
#endif
constexpr int kCWarps = ECO_V2_CWARPS;            // consumer warps (4 per SM sub-partition)
constexpr int kCThreads = kCWarps * 32;           // 512
constexpr int kThreads = kCThreads + 32;          // + the producer warp
constexpr int kTP = kCThreads * 2;                // pixels per tile
constexpr int kStages = 5;
constexpr int kStageBytes = 6 * kTP * 4;          // 24 KB
constexpr int kSmemBytes = kStages * kStageBytes; // 120 KB of dynamic shared memory

// barrier over the consumer warps only (the producer warp runs free after the prologue)
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(kCThreads) : "memory"); }
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct TileRange {
    int tpp;         // tiles per plane
    int t_lo, t_hi;  // this CTA's tiles (global tile index = n * tpp + k)
};
__device__ __forceinline__ TileRange tile_range(const CompArgs& a, int tp = kTP) {
    TileRange r;
    r.tpp = (int)((a.HW + tp - 1) / tp);
    const int64_t total = (int64_t)a.N * r.tpp;
    r.t_lo = (int)(total * blockIdx.x / gridDim.x);
    r.t_hi = (int)(total * (blockIdx.x + 1) / gridDim.x);
    return r;
}

struct PipeSmem {
    unsigned long long full[kStages];
    unsigned long long empty[kStages];
};

__device__ __forceinline__ void pipe_init(PipeSmem& ps) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&ps.full[s]), 1);
            mbar_init(smem_u32(&ps.empty[s]), kCWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// producer: ONE lane of the producer warp.  `k0` = tiles this CTA has pushed through the ring before this call
// (the mbarrier phases run on across the two passes of the fused kernel).
__device__ __forceinline__ void produce_tiles(const CompArgs& a, const TileRange& tr, bool reverse, uint32_t stage_base,
                                              PipeSmem& ps, int k0) {
    const float* xb = reinterpret_cast<const float*>(a.x);
    const float* gb = reinterpret_cast<const float*>(a.g);
    const int ntiles = tr.t_hi - tr.t_lo;
    if (ntiles <= 0) return;
    int t = reverse ? tr.t_hi - 1 : tr.t_lo;
    int n = t / tr.tpp, kk = t - n * tr.tpp;
    for (int k = 0; k < ntiles; ++k) {
        const int kg = k0 + k;
        const int s = kg % kStages;
        const uint32_t full = smem_u32(&ps.full[s]), empty = smem_u32(&ps.empty[s]);
        if (kg >= kStages) mbar_wait(empty, ((kg / kStages) - 1) & 1);
        const int64_t p0 = (int64_t)kk * kTP;
        const int valid = (int)((a.HW - p0 < kTP) ? (a.HW - p0) : kTP);
        const uint32_t bytes = (uint32_t)valid * 4u;
        const uint32_t dst = stage_base + (uint32_t)s * kStageBytes;
        const float* xs = xb + n * a.x_sn + p0;
        const float* gs = gb + n * a.g_sn + p0;
        mbar_expect_tx(full, 6u * bytes);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bulk_g2s(dst + (uint32_t)c * (kTP * 4), xs + c * a.x_sc, bytes, full);
            bulk_g2s(dst + (uint32_t)(3 + c) * (kTP * 4), gs + c * a.g_sc, bytes, full);
        }
        if (reverse) { if (--kk < 0) { kk = tr.tpp - 1; --n; } }
        else { if (++kk == tr.tpp) { kk = 0; ++n; } }
    }
}

// consumer side of one tile: wait, copy this thread's pixel pair of all six planes to registers, release the stage
__device__ __forceinline__ void consume_tile(uint32_t my_base, PipeSmem& ps, int kg, int lane, f2 (&z)[3], f2 (&g)[3]) {
    const int s = kg % kStages;
    mbar_wait(smem_u32(&ps.full[s]), (kg / kStages) & 1);
    const uint32_t sb = my_base + (uint32_t)s * kStageBytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        z[c] = lds_f2(sb + (uint32_t)c * (kTP * 4));
        g[c] = lds_f2(sb + (uint32_t)(3 + c) * (kTP * 4));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&ps.empty[s]));
}

// ---------------------------------------------------------------------------------------------
// pass 1: scalar, all 55 sums of a pixel in one thread
// ---------------------------------------------------------------------------------------------
enum : int {
    F_G = 0,     // + c           sum g_c
    F_NB = 3,    //               sum_c (g_c^2 - g_c)^2: non-zero iff some label is not exactly 0 or 1 (-> slow pass)
    F_X = 4,     // + c
    F_XX = 7,    // + c
    F_GX = 10,   // + c
    F_PAIR = 13, // + 14 p + {0 GD, 1 DS, 2 M1, 3 M1G, 4 M2, 5 M2G, 6 M3, 7 M3G, 8 UU1, 9 GU1, 10 UU2, 11 GU2, 12 UU3, 13 GU3}
    F_NACC = 55
};
constexpr int kFlushTiles = 16;    // 32 pixels per fp32 accumulator between folds into fp64

struct StatsSmem {
    double warp_slots[kCWarps][64];
    double sums[64];
    double corr[15];
    bool flag;
};

// PROB: the inputs already are probabilities (ECO_C3_PROBS: the reference's own call order, F.sigmoid at
// ess/train_multiclass.py:134 before losses_fn)
template <bool PROB = false>
__device__ __forceinline__ void stats_pixel(float z0, float z1, float z2, float g0, float g1, float g2, float (&acc)[F_NACC]) {
    const float x[3] = {PROB ? z0 : sigmoid_fast(z0), PROB ? z1 : sigmoid_fast(z1), PROB ? z2 : sigmoid_fast(z2)};
    const float g[3] = {g0, g1, g2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        acc[F_G + c] += g[c];
        acc[F_X + c] += x[c];
        acc[F_XX + c] = fmaf(x[c], x[c], acc[F_XX + c]);
        acc[F_GX + c] = fmaf(g[c], x[c], acc[F_GX + c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float e = fmaf(g[c], g[c], -g[c]);
        acc[F_NB] = fmaf(e, e, acc[F_NB]);
    }
    const float hh[2] = {fmaf(x[0], -0.5f, 0.5f), fmaf(x[1], -0.5f, 0.5f)};
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        float* pa = &acc[F_PAIR + 14 * p];
        const float xi = x[i], xj = x[j], gi = g[i], gj = g[j], h = hh[i];
        const float d = fabsf(xi - xj);
        const float gd = fabsf(gi - gj);
        const float m1 = xi * xj;
        const float q = xi * d;
        const float m3 = xi * q;
        const float u1 = fmaf(xj, h, xi);
        const float u2 = fmaf(d, h, xi);
        const float u3 = fmaf(q, h, xi);
        pa[0] += gd;
        pa[1] += d;
        pa[2] += m1;
        pa[3] = fmaf(m1, gj, pa[3]);
        pa[4] += q;
        pa[5] = fmaf(q, gd, pa[5]);
        pa[6] += m3;
        pa[7] = fmaf(m3, gd, pa[7]);
        pa[8] = fmaf(u1, u1, pa[8]);
        pa[9] = fmaf(gi, u1, pa[9]);
        pa[10] = fmaf(u2, u2, pa[10]);
        pa[11] = fmaf(gi, u2, pa[11]);
        pa[12] = fmaf(u3, u3, pa[12]);
        pa[13] = fmaf(gi, u3, pa[13]);
    }
}

__device__ __forceinline__ bool flush_flat_acc(float (&acc)[F_NACC], double* warp_slot /* smem [64] */, int lane) {
    const bool nonbinary = acc[F_NB] != 0.f;
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (grp * 32 + i < F_NACC) ? acc[grp * 32 + i] : 0.f;
        const float tot = butterfly32(v, lane);
        warp_slot[grp * 32 + lane] += (double)tot;
    }
#pragma unroll
    for (int k = 0; k < F_NACC; ++k) acc[k] = 0.f;
    return nonbinary;
}

// rare path: exact corrections of the label-b sums of one pixel whose labels are not all 0/1 (double math, shared
// atomics).  Slots 3L + {0 sum(b^2 - b), 1 softplus, 2 focal} for L = g1, g2, gd01, gd02, gd12 (eco_composite.cu).
__device__ __noinline__ void label_corrections_v2(float g0, float g1, float g2, double* corr) {
    const float lb[5] = {g1, g2, fabsf(g0 - g1), fabsf(g0 - g2), fabsf(g1 - g2)};
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

__device__ __forceinline__ void stats_smem_init(StatsSmem& sm) {
    for (int i = threadIdx.x; i < kCWarps * 64; i += blockDim.x) (&sm.warp_slots[0][0])[i] = 0.0;
    if (threadIdx.x < 15) sm.corr[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) sm.flag = false;
}

constexpr int kMaxGrid = 192;   // CTAs of one cooperative launch (one per SM)

// All-reduce of `count` (<= 128 - slot0) doubles across the ranks of a sharded step over NVLink peer memory, in the
// style of NCCL's LL protocol: every value travels as ONE 16-byte store {lo32, epoch, hi32, epoch} straight into
// every peer's exchange buffer, and the receiver spins on its own buffer until both tags carry this step's epoch.
// No fences and no separate flags: the data validates itself, so the latency is one NVLink write.  Rows are double
// buffered by epoch parity (a rank can be at most one exchange ahead of a peer that has not read yet); the region
// starts `xch_ll_offset_bytes(world)` into the buffer, after the first-generation slots.  Every rank adds the `world`
// contributions in rank order, so all ranks hold bit-identical totals.  Called by all CONSUMER threads of ONE CTA.
__host__ __device__ inline size_t xch_ll_offset_bytes(int world) { return xch_flags_offset_doubles(world) * 8 + 512; }
// Deterministic grid-wide sums without a serial "last CTA adds everything" phase: every CTA adds its partial into
// two 64-bit INTEGER accumulators per sum (v * 2^30 rounded to an integer, and the rounding remainder * 2^32), so the
// total does not depend on the order of the atomics (integer addition is associative) and is exact to 2^-62.
// Range: |sum| < 2^33 (the host routes larger problems to the first-generation kernels).
constexpr double kFixHi = 1073741824.0;            // 2^30
constexpr double kFixLo = 4294967296.0;            // 2^32
__device__ __forceinline__ void fix_add(unsigned long long* slot2, double v) {
    const double sc = v * kFixHi;
    const long long hi = __double2ll_rn(sc);
    const long long lo = __double2ll_rn((sc - (double)hi) * kFixLo);
    atomicAdd(slot2, (unsigned long long)hi);
    atomicAdd(slot2 + 1, (unsigned long long)lo);
}
// the accumulators are replicated kFixRep times (CTA b adds into replica b % kFixRep) to spread the L2 atomics
constexpr int kFixRep = 8;
__device__ __forceinline__ double fix_get(const unsigned long long* slot2, int rep_stride) {
    unsigned long long hi = 0ull, lo = 0ull;
#pragma unroll
    for (int r = 0; r < kFixRep; ++r) {
        hi += __ldcg(slot2 + (size_t)r * rep_stride);
        lo += __ldcg(slot2 + (size_t)r * rep_stride + 1);
    }
    return ((double)(long long)hi + (double)(long long)lo * (1.0 / kFixLo)) * (1.0 / kFixHi);
}

// workspace words of multiclass3_fused_v2_kernel (all zero between launches)
// fix1 is double buffered by step parity: step k adds into buffer k & 1 while CTA 0 clears buffer (k + 1) & 1 for the
// next step, so nobody waits for the clearing.  tr2[0] counts the CTAs that have finished pass 2.
struct V2Ws {
    unsigned int arrive1, step, _pad0, _pad1;
    unsigned long long tr2[2];
    unsigned long long fix1[2][kFixRep][2 * kNAcc];   // pass-1 sums
};

// ---------------------------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------------------------
// per-leaf scalars in shared memory.  U-type leaves (a = label, b real): 0..2 = channel leaves, 3 + 3p + k = U_{k+1}
// of pair p.  I-type leaves (a = product, b = label): 3p + k = I_{k+1} of pair p.
struct Coef2 {
    float4 ua[12];  // {c_Sb + c_SP/2, c_Sab, 2 c_Sbb, c_SP}
    float4 uw[12];  // {R0 w, R1 w, R2 w, w or |w|^(2/3)}: softplus-remainder polynomial times the leaf scale w; focal weight
    float ufl[12];  // c_FL
    float2 ia[9];   // {c_Sa, c_Sab}
};

// weights of the linear sums of the real-b leaves (needed by the pre-pass, before the coefficients exist).
// pos = every such leaf has a scale >= 0 (always true for the reference's weights): the focal weight is then stored as
// w^(2/3), which rides for free on 1 - b:  (w^(2/3) (1-b))^1.5 = w (1-b)^1.5.
__device__ __forceinline__ int u_leaf_of(int t) {
    if (t < 3) return t;
    return ((t - 3) % 6 & 1) ? 3 + 3 * ((t - 3) / 6) + (((t - 3) % 6) >> 1) : -1;
}
__device__ __forceinline__ void fill_weights(Coef2& c2, const double* scale, int t, bool pos) {
    if (t >= ECO_C3_NLEAF) return;
    const int ul = u_leaf_of(t);
    if (ul < 0) return;
    const double w = scale[t];
    c2.uw[ul] = make_float4((float)(w * (double)kSpR0), (float)(w * (double)kSpR1), (float)(w * (double)kSpR2),
                            pos ? (float)cbrt(w * w) : (float)w);
}

__device__ __forceinline__ void fill_coef2(Coef2& c2, const LeafCoef* cf, int t) {
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
        c2.ua[ul] = make_float4(c.sb + 0.5f * c.sp, c.sab, c.sbb2, c.sp);
        c2.ufl[ul] = c.fl;
    } else {
        c2.ia[il] = make_float2(c.sa, c.sab);
    }
}

__device__ __forceinline__ float xor_sign(float a, float s) {
    return __uint_as_float(__float_as_uint(a) ^ (__float_as_uint(s) & 0x80000000u));
}
__device__ __forceinline__ f2 apply_sign(f2 v, f2 s) { return make_float2(xor_sign(v.x, s.x), xor_sign(v.y, s.y)); }
__device__ __forceinline__ f2 neg2(f2 v) { return make_float2(-v.x, -v.y); }

// the linear sums of a real-b leaf: weighted softplus remainder t^2 r(t) (t = b^2) and weighted focal term in log2
// units.  One definition for pass 2 and for the pre-pass below: the same operations in the same order, bit for bit.
__device__ __forceinline__ void leaf_tr_sp(const float4 cw, f2 t, f2& sp_acc) {
    f2 q = fma2(t, splat(cw.z), splat(cw.y));   // w r(t), the leaf scale folded into the coefficients
    q = fma2(q, t, splat(cw.x));
    sp_acc = fma2(mul2(t, t), q, sp_acc);
}
// focal term given om = 1 - b (POSW: times w^(2/3)), sq = sqrt(om), lg = lg2(b + eps)
template <bool POSW>
__device__ __forceinline__ void leaf_tr_fl(const float4 cw, f2 om, f2 sq, f2 lg, f2& fl_acc) {
    if (POSW) fl_acc = fma2(mul2(om, sq), lg, fl_acc);
    else fl_acc = fma2(mul2(mul2(om, sq), lg), splat(cw.w), fl_acc);
}
template <bool POSW>
__device__ __forceinline__ void leaf_tr(const float4 cw, f2 b, f2 t, f2& sp_acc, f2& fl_acc) {
    leaf_tr_sp(cw, t, sp_acc);
    const f2 om = POSW ? fma2(b, splat(-cw.w), splat(cw.w)) : fma2(b, splat(-1.0f), splat(1.0f));
    const f2 sq = make_float2(sqrt_approx(om.x), sqrt_approx(om.y));
    const f2 be = add2(b, splat(kEps));
    const f2 lg = make_float2(lg2_approx(be.x), lg2_approx(be.y));
    leaf_tr_fl<POSW>(cw, om, sq, lg, fl_acc);
}

// dT/db of a U-type leaf at b (a = label); SIG: the BCE term carries gradient; FL: the focal term carries gradient
template <bool SIG, bool FL>
__device__ __forceinline__ f2 leaf_g(const float4 ca, const float cfl, f2 a, f2 b) {
    f2 t;
    if (SIG) t = mul2(b, b);
    const f2 k = fma2(a, splat(ca.y), splat(ca.x));   // c_Sab a + c0'
    f2 r;
    if (SIG) {
        f2 s = fma2(t, splat(kSgS3), splat(kSgS2));
        s = fma2(s, t, splat(kSgS1));
        s = fma2(s, t, splat(kSgS0));
        const f2 w = fma2(s, splat(ca.w), splat(ca.z));
        r = fma2(b, w, k);
    } else {
        r = fma2(b, splat(ca.z), k);
    }
    if (FL) {
        const f2 om = fma2(b, splat(-1.0f), splat(1.0f));
        const f2 sq = make_float2(sqrt_approx(om.x), sqrt_approx(om.y));
        const f2 be = add2(b, splat(kEps));
        const f2 lg = make_float2(lg2_approx(be.x), lg2_approx(be.y));
        const f2 w15 = mul2(om, sq);
        // + c_FL d/db[-(1-b)^1.5 log(b+eps)] = c_FL (1.5 ln2 sqrt(1-b) lg2(b+eps) - (1-b)^1.5 / (b+eps))
        const f2 rc = make_float2(rcp_approx(be.x), rcp_approx(be.y));
        const f2 v = fma2(mul2(sq, splat(1.5f * kLn2)), lg, neg2(mul2(w15, rc)));
        r = fma2(v, splat(cfl), r);
    }
    return r;
}

// the whole gradient of one pixel pair: x = probabilities, g = labels, diffs[p] = x_i - x_j
template <bool SIG, bool FL>
__device__ __forceinline__ void pixel_pair_grad2(const f2 (&x)[3], const f2 (&g)[3], const f2 (&diffs)[3],
                                                 const Coef2& c2, f2 (&gx)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gx[c] = leaf_g<SIG, FL>(c2.ua[c], c2.ufl[c], g[c], x[c]);
    f2 hh[2];
    hh[0] = fma2(x[0], splat(-0.5f), splat(0.5f));
    hh[1] = fma2(x[1], splat(-0.5f), splat(0.5f));
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        const f2 xi = x[i], xj = x[j], gi = g[i], gj = g[j], h = hh[i];
        const f2 diff = diffs[p];
        const f2 d = abs2(diff);
        const f2 gd = abs2(add2(gi, neg2(gj)));
        const f2 q = mul2(xi, d);                 // a2 = x_i d   (a1 = x_i x_j and a3 = x_i q are not needed themselves)
        const f2 u1 = fma2(xj, h, xi);
        const f2 u2 = fma2(d, h, xi);
        const f2 u3 = fma2(q, h, xi);
        const f2 G1 = leaf_g<SIG, FL>(c2.ua[3 + 3 * p + 0], c2.ufl[3 + 3 * p + 0], gi, u1);
        const f2 G2 = leaf_g<SIG, FL>(c2.ua[3 + 3 * p + 1], c2.ufl[3 + 3 * p + 1], gi, u2);
        const f2 G3 = leaf_g<SIG, FL>(c2.ua[3 + 3 * p + 2], c2.ufl[3 + 3 * p + 2], gi, u3);
        const float2 k1 = c2.ia[3 * p + 0], k2 = c2.ia[3 * p + 1], k3 = c2.ia[3 * p + 2];
        const f2 A1 = fma2(gj, splat(k1.y), splat(k1.x));
        const f2 A2 = fma2(gd, splat(k2.y), splat(k2.x));
        const f2 A3 = fma2(gd, splat(k3.y), splat(k3.x));
        // totals w.r.t. the intermediates: Q = dT/dq, D = dT/dd
        f2 Q = fma2(A3, xi, A2);
        Q = fma2(G3, h, Q);
        f2 D = mul2(Q, xi);
        D = fma2(G2, h, D);
        const f2 Ds = apply_sign(D, diff);        // dT/d(x_i - x_j) through the |.| kink
        // x_i: direct terms of u_k (1 - p_k/2), a1 (x_j), q (d), a3 (q), plus the kink.  (Grouping by multiplier,
        // x_j (A1 - G1/2) + d (Q - G2/2) + q (A3 - G3/2), is 4 issue cycles shorter on paper and 2 us slower measured.)
        f2 E = mul2(G1, xj);
        E = fma2(G2, d, E);
        E = fma2(G3, q, E);
        f2 gi_acc = add2(add2(G1, G2), add2(G3, gx[i]));
        gi_acc = fma2(E, splat(-0.5f), gi_acc);
        gi_acc = fma2(A1, xj, gi_acc);
        gi_acc = fma2(Q, d, gi_acc);
        gi_acc = fma2(A3, q, gi_acc);
        gx[i] = add2(gi_acc, Ds);
        f2 gj_acc = fma2(A1, xi, gx[j]);
        gj_acc = fma2(G1, h, gj_acc);
        gx[j] = add2(gj_acc, neg2(Ds));
    }
}

// rare: a pixel whose probabilities tie to within kTieEps -- the sign of the |x_i - x_j| kink (and sign(0) = 0)
// must come from ATen's exact sigmoid bits.  Scalar path of the first-generation kernel.
// With `prob` the inputs are the probabilities themselves and the gradient is taken w.r.t. them.
__device__ __noinline__ float3 tie_pixel_grad(float z0, float z1, float z2, float g0, float g1, float g2,
                                              const LeafCoef* cf, bool need_sig, bool need_fl, bool prob = false) {
    const float x[3] = {prob ? z0 : sigmoid_exact(z0), prob ? z1 : sigmoid_exact(z1), prob ? z2 : sigmoid_exact(z2)};
    const float g[3] = {g0, g1, g2};
    float gx[3];
    pixel_grad(x, g, cf, need_sig, need_fl, gx);
    // returned BY VALUE: taking the address of the caller's outputs would push them through local memory on every tile
    if (prob) return make_float3(gx[0], gx[1], gx[2]);
    return make_float3(gx[0] * ((1.0f - x[0]) * x[0]), gx[1] * ((1.0f - x[1]) * x[1]), gx[2] * ((1.0f - x[2]) * x[2]));
}

// only the linear sums of one pixel pair, in the leaf order of pixel_pair_grad2
template <bool POSW>
__device__ __forceinline__ void pixel_pair_tr(const f2 (&x)[3], const Coef2& c2, f2& sp_acc, f2& fl_acc) {
#pragma unroll
    for (int c = 0; c < 3; ++c) leaf_tr<POSW>(c2.uw[c], x[c], mul2(x[c], x[c]), sp_acc, fl_acc);
    f2 hh[2];
    hh[0] = fma2(x[0], splat(-0.5f), splat(0.5f));
    hh[1] = fma2(x[1], splat(-0.5f), splat(0.5f));
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int i = pair_i(p), j = pair_j(p);
        const f2 xi = x[i], xj = x[j], h = hh[i];
        const f2 d = abs2(add2(xi, neg2(xj)));
        const f2 q = mul2(xi, d);
        const f2 us[3] = {fma2(xj, h, xi), fma2(d, h, xi), fma2(q, h, xi)};
#pragma unroll
        for (int k = 0; k < 3; ++k) leaf_tr<POSW>(c2.uw[3 + 3 * p + k], us[k], mul2(us[k], us[k]), sp_acc, fl_acc);
    }
}

// consumer side of pass 2 over this CTA's tiles, forwards (the stand-alone gradient kernel)
template <bool SIG, bool FL>
__device__ __forceinline__ void grad_consume(const CompGradArgs& ga, const TileRange& tr, uint32_t stage_base, PipeSmem& ps,
                                             const Coef2& c2, const LeafCoef* cf) {
    const CompArgs& a = ga.a;
    float* __restrict__ ob = reinterpret_cast<float*>(ga.gx);
    const int lane = threadIdx.x & 31;
    const int ntiles = tr.t_hi - tr.t_lo;
    if (ntiles <= 0) return;
    int n = tr.t_lo / tr.tpp, kk = tr.t_lo - n * tr.tpp;
    const uint32_t my = stage_base + threadIdx.x * 8;
    const int pix = 2 * (int)threadIdx.x;
    for (int k = 0; k < ntiles; ++k) {
        f2 z[3], g[3];
        consume_tile(my, ps, k, lane, z, g);
        const int64_t p0 = (int64_t)kk * kTP;
        if (p0 + pix < a.HW) {
            f2 x[3], gx[3], diffs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) x[c] = sigmoid_fast2(z[c]);
#pragma unroll
            for (int p = 0; p < 3; ++p) diffs[p] = add2(x[pair_i(p)], neg2(x[pair_j(p)]));
            pixel_pair_grad2<SIG, FL>(x, g, diffs, c2, gx);
            f2 o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = mul2(gx[c], mul2(x[c], fma2(x[c], splat(-1.0f), splat(1.0f))));
            const float dx = fminf(fminf(fabsf(diffs[0].x), fabsf(diffs[1].x)), fabsf(diffs[2].x));
            const float dy = fminf(fminf(fabsf(diffs[0].y), fabsf(diffs[1].y)), fabsf(diffs[2].y));
            if (fminf(dx, dy) < kTieEps) {
                if (dx < kTieEps) {
                    const float3 r = tie_pixel_grad(z[0].x, z[1].x, z[2].x, g[0].x, g[1].x, g[2].x, cf, SIG, FL);
                    o[0].x = r.x; o[1].x = r.y; o[2].x = r.z;
                }
                if (dy < kTieEps) {
                    const float3 r = tie_pixel_grad(z[0].y, z[1].y, z[2].y, g[0].y, g[1].y, g[2].y, cf, SIG, FL);
                    o[0].y = r.x; o[1].y = r.y; o[2].y = r.z;
                }
            }
            float* op = ob + n * ga.gx_sn + p0 + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) stg_stream_f2(op + c * ga.gx_sc, o[c]);
        }
        if (++kk == tr.tpp) { kk = 0; ++n; }
    }
}

// stand-alone pass 2 on the tile pipeline: same contract as composite3_grad_packed_kernel (fp32, from logits)
__global__ void __launch_bounds__(kThreads, 1)
composite3_grad_v2_kernel(CompGradArgs ga, const double* __restrict__ jac, const float* __restrict__ upstream) {
    extern __shared__ __align__(128) char stage_smem[];
    __shared__ LeafCoef cf[ECO_C3_NLEAF];
    __shared__ Coef2 c2;
    __shared__ PipeSmem ps;
    if (threadIdx.x < ECO_C3_NLEAF) cf[threadIdx.x] = make_coef(jac + threadIdx.x * ECO_NLOSS * ECO_NJAC, upstream);
    pipe_init(ps);
    fill_coef2(c2, cf, threadIdx.x);
    __syncthreads();
    const TileRange tr = tile_range(ga.a);
    const uint32_t sbase = smem_u32(stage_smem);
    if (threadIdx.x >= kCThreads) {
        if (threadIdx.x == kCThreads) produce_tiles(ga.a, tr, false, sbase, ps, 0);
    } else {
        // block-uniform dispatch on which of the 7 outputs carry gradient (train_multiclass.py:145 weights them 0 / 1)
        const bool need_sig = upstream[1] != 0.f, need_fl = upstream[2] != 0.f;
        if (need_fl) {
            if (need_sig) grad_consume<true, true>(ga, tr, sbase, ps, c2, cf);
            else grad_consume<false, true>(ga, tr, sbase, ps, c2, cf);
        } else {
            if (need_sig) grad_consume<true, false>(ga, tr, sbase, ps, c2, cf);
            else grad_consume<false, false>(ga, tr, sbase, ps, c2, cf);
        }
    }
}

}  // namespace v2
}  // namespace eco
