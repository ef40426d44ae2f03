// One-launch step of ONE leaf whose whole input fits the shared memory of the GPU (BASELINE configs[0], cfg1:
// ORGANS=whole_body, 54 x 1 x 256 x 256 fp32 = 28.3 MB of logits + masks against 148 x 227 KB = 33.6 MB of shared memory).
// Included by eco_leaf.cu (it uses PairArgs / FusedArgs / LeafCoef from there); selected by eco_pair_fused.
//
// Replaces, for that configuration, `losses_fn` + `loss.backward()` of ess/train_multiclass.py:139-147 (C == 1:
// ess/loss_composite.py:32-40 / ess/train_multiclass.py:264-274, prediction in the gt slot, background_weight honoured).
//
// pair_fused_kernel reads its input twice (pass 1 from HBM, pass 2 from L2) and hands the sums over through the channel's
// last CTA.  Here every CTA pulls its contiguous share of both tensors into shared memory ONCE (1-D TMA bulk copies in
// chunks, one mbarrier per chunk, all issued up front), takes the statistics chunk by chunk as the data lands, writes the
// probabilities back into the resident tile (pass 2 needs no second sigmoid), adds its partial sums into 64-bit INTEGER
// accumulators (exact, order-independent: deterministic without a serial last-CTA reduction), and after the grid-wide
// arrival every CTA derives the closed forms itself and streams the gradient out of shared memory: the only DRAM traffic
// is the algorithmic 12 B/element, and nothing between the passes waits for a memory round trip beyond the arrival itself.
#pragma once

namespace eco {
namespace resident {

constexpr int kRThreads = 512;
constexpr int kRChunks = 8;                       // TMA chunks per tensor and CTA
constexpr int kRMaxSmem = 220 * 1024;             // dynamic shared memory budget for the two resident tiles
constexpr int kRRep = 8;                          // replicas of the integer accumulators (spreads the L2 atomics)
constexpr double kRFixHi = 1073741824.0;          // 2^30
constexpr double kRFixLo = 4294967296.0;          // 2^32

__device__ __forceinline__ uint32_t rs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void rs_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// workspace (all zero before the first launch), double buffered by step parity like the composite kernel's
struct ResidentWs {
    unsigned int step, _pad[3];
    unsigned int arrive[2][4];
    unsigned long long fix[2][kRRep][16];    // 7 sums x (hi, lo)
};

__device__ __forceinline__ void rs_fix_add(unsigned long long* slot2, double v) {
    const double sc = v * kRFixHi;
    const long long hi = __double2ll_rn(sc);
    const long long lo = __double2ll_rn((sc - (double)hi) * kRFixLo);
    atomicAdd(slot2, (unsigned long long)hi);
    atomicAdd(slot2 + 1, (unsigned long long)lo);
}
__device__ __forceinline__ double rs_fix_get(const unsigned long long* slot2) {
    unsigned long long hi = 0ull, lo = 0ull;
#pragma unroll
    for (int r = 0; r < kRRep; ++r) {
        hi += __ldcg(slot2 + (size_t)r * 16);
        lo += __ldcg(slot2 + (size_t)r * 16 + 1);
    }
    return ((double)(long long)hi + (double)(long long)lo * (1.0 / kRFixLo)) * (1.0 / kRFixHi);
}

struct ResidentSmem {
    unsigned long long bar[kRChunks];
    double warp_sums[kRThreads / 32][7];
    double sums[ECO_NSTAT];
    double sl[ECO_NLOSS];
    double sj[ECO_NLOSS][ECO_NJAC];
    LeafCoef cf;
    float up[ECO_NLOSS + 1];
};

// a, b: fully contiguous fp32 tensors of `total` elements (a multiple of 4); CTA i owns the 4-element units
// [units * i / grid, units * (i + 1) / grid).
__global__ void __launch_bounds__(kRThreads, 1)
leaf_resident_kernel(FusedArgs fa, int64_t total, const float* __restrict__ upstream, ResidentWs* __restrict__ ws,
                     double* __restrict__ sums_out, float* __restrict__ losses_out) {
    extern __shared__ __align__(128) char tile_smem[];
    __shared__ ResidentSmem rs;
    if (upstream_unchanged(upstream, fa.upstream_prev)) return;
    const PairArgs& p = fa.g.p;
    const bool a_logit = p.flags & ECO_A_LOGIT, b_logit = p.flags & ECO_B_LOGIT, need_bg = p.flags & ECO_NEED_BG;
    const int64_t units = total / 4;
    const int64_t u_lo = units * blockIdx.x / gridDim.x, u_hi = units * (blockIdx.x + 1) / gridDim.x;
    const int cnt = (int)(u_hi - u_lo) * 4;                       // this CTA's elements
    const int per_chunk = ((cnt / 4 + kRChunks - 1) / kRChunks) * 4;   // elements per chunk (a multiple of 4)
    float* a_s = reinterpret_cast<float*>(tile_smem);
    float* b_s = a_s + ((cnt + 31) / 32) * 32;
    const float* a_g = reinterpret_cast<const float*>(p.a) + u_lo * 4;
    const float* b_g = reinterpret_cast<const float*>(p.b) + u_lo * 4;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kRChunks; ++k) rs_mbar_init(rs_smem_u32(&rs.bar[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the whole share of both tensors, issued up front: nothing in front of the first TMA waits on global memory
        for (int k = 0; k < kRChunks; ++k) {
            const int e0 = k * per_chunk;
            const int n = cnt - e0 < per_chunk ? cnt - e0 : per_chunk;
            const uint32_t bar = rs_smem_u32(&rs.bar[k]);
            if (n <= 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); continue; }
            rs_mbar_expect_tx(bar, 2u * (uint32_t)n * 4u);
            rs_bulk_g2s(rs_smem_u32(a_s + e0), a_g + e0, (uint32_t)n * 4u, bar);
            rs_bulk_g2s(rs_smem_u32(b_s + e0), b_g + e0, (uint32_t)n * 4u, bar);
        }
    }
    const int par = (int)(__ldcg(&ws->step) & 1u);
    if (blockIdx.x == 0) {   // clear the other parity's buffers for the next step
        if (threadIdx.x < kRRep * 16) (&ws->fix[par ^ 1][0][0])[threadIdx.x] = 0ull;
        if (threadIdx.x == 0) ws->arrive[par ^ 1][0] = 0u;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + ECO_NLOSS) rs.up[threadIdx.x - 32] = upstream[threadIdx.x - 32];
    __syncthreads();   // barriers initialised

    // ---- pass 1: statistics chunk by chunk as the data lands; probabilities go back into the resident tile ----------------
    float acc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = 0.f;
    double dacc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) dacc[k] = 0.0;
    for (int k = 0; k < kRChunks; ++k) {
        rs_mbar_wait(rs_smem_u32(&rs.bar[k]), 0);
        const int e0 = k * per_chunk;
        const int e1 = e0 + per_chunk < cnt ? e0 + per_chunk : cnt;
        for (int e = e0 + (int)threadIdx.x * 4; e < e1; e += kRThreads * 4) {
            float4 av = *reinterpret_cast<const float4*>(a_s + e);
            float4 bv = *reinterpret_cast<const float4*>(b_s + e);
            float* ap = reinterpret_cast<float*>(&av);
            float* bp = reinterpret_cast<float*>(&bv);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float a = ap[v], b = bp[v];
                if (a_logit) { a = sigmoid_fast(a); ap[v] = a; }
                if (b_logit) { b = sigmoid_fast(b); bp[v] = b; }
                acc[0] += a;
                acc[1] += b;
                acc[2] = fmaf(a, b, acc[2]);
                acc[3] = fmaf(b, b, acc[3]);
                acc[4] += fmaf(softplus_neg_abs_log2(b), kLn2, fmaxf(b, 0.f));   // (MUFU form: this kernel is a latency chain, not XU-limited -- the polynomial of the pair-leaf kernels costs it 1 us at cfg1)
                acc[5] -= focal_fg_log2(b);
                if (need_bg) acc[6] -= focal_bg_log2(b);
            }
            if (a_logit) *reinterpret_cast<float4*>(a_s + e) = av;
            if (b_logit) *reinterpret_cast<float4*>(b_s + e) = bv;
        }
        // <= 64 values per fp32 partial: fold after every chunk (a thread sees per_chunk / kRThreads elements of it)
#pragma unroll
        for (int j = 0; j < 7; ++j) { dacc[j] += (double)acc[j]; acc[j] = 0.f; }
    }
    dacc[5] *= kLn2d;   // the focal terms were accumulated in log2 units
    dacc[6] *= kLn2d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const double v = warp_sum(dacc[j]);
        if (lane == 0) rs.warp_sums[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kRThreads / 32; ++w) v += rs.warp_sums[w][threadIdx.x];
        rs_fix_add(ws->fix[par][blockIdx.x % kRRep] + 2 * threadIdx.x, v);
    }
    __syncthreads();
    // ---- grid-wide hand-over: arrive, wait for everybody, read the totals ---------------------------------------------------
    if (threadIdx.x == 0) {
        __threadfence();   // cumulative over the barrier: the seven threads' atomics precede the arrival
        atomicAdd(&ws->arrive[par][0], 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&ws->arrive[par][0]) : "memory");
            if (seen < gridDim.x) __nanosleep(32);
        } while (seen < gridDim.x);
    }
    __syncthreads();
    if (threadIdx.x < 7) rs.sums[1 + threadIdx.x] = rs_fix_get(ws->fix[par][0] + 2 * threadIdx.x);
    if (threadIdx.x == 7) rs.sums[S_N] = (double)total;
    __syncthreads();
    // ---- closed forms, redundantly per CTA: one warp per loss kind (seven different float64 code paths) ----------------
    if (lane == 0 && warp < ECO_NLOSS) {
        if (fa.shaped) leaf_closed_form_row<false>(rs.sums, fa.bw, fa.scale, warp, rs.sl[warp], rs.sj[warp], fa.shape);
        else leaf_closed_form_row(rs.sums, fa.bw, fa.scale, warp, rs.sl[warp], rs.sj[warp]);
    }
    __syncthreads();
    if (threadIdx.x < ECO_NJAC) {
        const int j = threadIdx.x;
        double v = 0.0;
#pragma unroll
        for (int k = 1; k < ECO_NLOSS; ++k) {
            const double w = (double)rs.up[k];
            if (w != 0.0) v += w * rs.sj[k][j];   // rows without upstream weight stay out (their Jacobian may be non-finite)
        }
        reinterpret_cast<float*>(&rs.cf)[j] = (float)(j == 3 ? 2.0 * v : v);
        if (blockIdx.x == 0) losses_out[j] = (float)rs.sl[j];
    }
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + ECO_NSTAT) sums_out[threadIdx.x - 32] = rs.sums[threadIdx.x - 32];
    __syncthreads();

    // ---- pass 2: the gradient straight out of the resident tiles -----------------------------------------------------------
    const LeafCoef cf = rs.cf;
    const bool need_sig = cf.sp != 0.f, need_fl = cf.fl != 0.f, need_flb = cf.flb != 0.f;
    float* ga_g = fa.g.ga ? reinterpret_cast<float*>(fa.g.ga) + u_lo * 4 : nullptr;
    float* gb_g = fa.g.gb ? reinterpret_cast<float*>(fa.g.gb) + u_lo * 4 : nullptr;
    for (int e = (int)threadIdx.x * 4; e < cnt; e += kRThreads * 4) {
        const float4 av = *reinterpret_cast<const float4*>(a_s + e);
        const float4 bv = *reinterpret_cast<const float4*>(b_s + e);
        const float* ap = reinterpret_cast<const float*>(&av);
        const float* bp = reinterpret_cast<const float*>(&bv);
        float oa[4], ob[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float a = ap[v], b = bp[v];   // probabilities where the slot held logits
            float da = fmaf(cf.sab, b, cf.sa);
            float db = fmaf(cf.sab, a, fmaf(cf.sbb2, b, cf.sb));
            if (need_sig) db = fmaf(cf.sp, sigmoid_fast(b), db);
            if (need_fl) db = fmaf(cf.fl, dfocal_fg(b), db);
            if (need_flb) db = fmaf(cf.flb, dfocal_bg(b), db);
            if (a_logit) da *= (1.0f - a) * a;
            if (b_logit) db *= (1.0f - b) * b;
            oa[v] = da;
            ob[v] = db;
        }
        if (ga_g) stg_stream_f4(ga_g + e, make_float4(oa[0], oa[1], oa[2], oa[3]));
        if (gb_g) stg_stream_f4(gb_g + e, make_float4(ob[0], ob[1], ob[2], ob[3]));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ws->step = (unsigned int)par + 1u;   // every CTA read `step` before it arrived
}

// bytes of dynamic shared memory one CTA needs for `total` elements over `grid` CTAs; 0 = does not fit
static inline int resident_smem_bytes(int64_t total, int grid) {
    const int64_t units = total / 4;
    const int64_t per = (units + grid - 1) / grid * 4;            // elements of the largest share
    const int64_t padded = (per + 31) / 32 * 32;
    const int64_t bytes = 2 * padded * 4 + 128;
    return bytes <= kRMaxSmem ? (int)bytes : 0;
}

}  // namespace resident
}  // namespace eco
