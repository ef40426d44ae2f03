// Fused plain 3-organ multi-class loss step: the loss train_multiclass.py actually trains with.
//   ess/train_multiclass.py:253-274 `losses_fn(outputs, labels, composite_set_theory=False, ...)` for C > 1 is the sum
//   over channels of the 7-loss leaf (a = label g_c, b = prediction x_c) (:260-262; loss_composite.py:28-40 is the same
//   with every leaf doubled), preceded by `F.sigmoid` (:134) and followed by `loss.backward()` (:147).
// One cooperative launch on the tile pipeline of eco_composite_v2.cuh: pass 1 (18 sums per pixel) -> integer grid sums
// -> closed forms per CTA -> pass 2 (gradient w.r.t. the logits).  Unlike the composite loss this IS a streaming
// problem: ~50 packed instructions per pixel pair and pass, so the TMA ring, not instruction issue, sets the pace.
// fp32 logits, 16-byte aligned planes, C == 3 (the ring carries exactly the six planes x0 x1 x2 g0 g1 g2).
#pragma once

namespace eco {
namespace v2 {

enum : int { M_G = 0, M_X, M_XX, M_GX, M_R, M_FL, M_PER = 6, M_NACC = 18 };
constexpr int kMcSums = 1 + 3 * M_PER;   // n, then per channel {sum g, sum x, sum x^2, sum g x, SP, FL}

struct McSmem {
    double warp_slots[kCWarps][32];
    double sums[32];
    double acc[kMcSums];
    double sl[3][ECO_NLOSS];
    double jac_s[3][ECO_NLOSS][ECO_NJAC];
    LeafCoef cf[3];
    float4 ua[3];
    float ufl[3];
    float up[ECO_NLOSS + 1];
    bool flag;
    PipeSmem ps;
};

__device__ __forceinline__ void mc_flush(f2 (&acc)[M_NACC], double* warp_slot /* smem [32] */, int lane) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < M_NACC) ? acc[i].x + acc[i].y : 0.f;
    const float tot = butterfly32(v, lane);
    warp_slot[lane] += (double)tot;
#pragma unroll
    for (int k = 0; k < M_NACC; ++k) acc[k] = splat(0.f);
}

// gradient-side leaf: dT/db at b = x_c with a = g_c (leaf_g of the composite kernel)
template <bool SIG, bool FL>
__device__ __forceinline__ f2 mc_leaf_g(const float4 ca, const float cfl, f2 a, f2 b) { return leaf_g<SIG, FL>(ca, cfl, a, b); }

// PROB: the inputs already are probabilities (the reference's own call order, F.sigmoid at ess/train_multiclass.py:134
// before losses_fn) and the gradient is taken w.r.t. them; torch's sigmoid backward carries it on to the logits
template <bool SIG, bool FL, bool PROB>
__device__ __forceinline__ void mc_grad_consume(const CompGradArgs& ga, const TileRange& tr, uint32_t stage_base, McSmem& ms,
                                                int k0) {
    const CompArgs& a = ga.a;
    float* __restrict__ ob = reinterpret_cast<float*>(ga.gx);
    const int lane = threadIdx.x & 31;
    const int ntiles = tr.t_hi - tr.t_lo;
    int t = tr.t_hi - 1;
    int n = t / tr.tpp, kk = t - n * tr.tpp;
    const uint32_t my = stage_base + threadIdx.x * 8;
    const int pix = 2 * (int)threadIdx.x;
    for (int k = 0; k < ntiles; ++k) {
        f2 z[3], g[3];
        consume_tile(my, ms.ps, k0 + k, lane, z, g);
        const int64_t p0 = (int64_t)kk * kTP;
        if (p0 + pix < a.HW) {
            float* op = ob + n * ga.gx_sn + p0 + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const f2 x = PROB ? z[c] : sigmoid_fast2(z[c]);
                const f2 G = mc_leaf_g<SIG, FL>(ms.ua[c], ms.ufl[c], g[c], x);
                stg_stream_f2(op + c * ga.gx_sc, PROB ? G : mul2(G, mul2(x, fma2(x, splat(-1.0f), splat(1.0f)))));
            }
        }
        if (--kk < 0) { kk = tr.tpp - 1; --n; }
    }
}

// upstream_prev != nullptr: the "only if changed" form of the drop-in autograd path -- the outputs already hold the step for
// the weights `upstream_prev`; if `upstream` is the same vector the whole grid leaves before it touches anything.
template <bool PROB>
__global__ void __launch_bounds__(kThreads, 1)
multiclass3_fused_v2_kernel(CompGradArgs ga, double scale, const float* __restrict__ upstream, V2Ws* __restrict__ ws,
                            float* __restrict__ losses_out, const float* __restrict__ upstream_prev) {
    extern __shared__ __align__(128) char stage_smem[];
    __shared__ McSmem ms;
    if (upstream_prev) {
        bool same = true;
#pragma unroll
        for (int k = 0; k < ECO_NLOSS; ++k) same = same && (__float_as_uint(upstream[k]) == __float_as_uint(upstream_prev[k]));
        if (same) return;
    }
    for (int i = threadIdx.x; i < kCWarps * 32; i += kThreads) (&ms.warp_slots[0][0])[i] = 0.0;
    pipe_init(ms.ps);
    const CompArgs& a = ga.a;
    const TileRange tr = tile_range(a);
    const int ntiles = tr.t_hi - tr.t_lo;
    const uint32_t sbase = smem_u32(stage_smem);
    if (threadIdx.x >= kCThreads) {
        if (threadIdx.x == kCThreads) {
            produce_tiles(a, tr, false, sbase, ms.ps, 0);
            produce_tiles(a, tr, true, sbase, ms.ps, ntiles);
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ECO_TL(0);
    // the workspace is double buffered by step parity (integer accumulators in fix1, arrival counter in tr2): step k works
    // in buffer k & 1 while CTA 0 clears buffer (k + 1) & 1, so there is no re-arming phase at the end of a step
    const int par = (int)(__ldcg(&ws->step) & 1u);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < kFixRep * 2 * kNAcc; i += kCThreads) (&ws->fix1[par ^ 1][0][0])[i] = 0ull;
        if (threadIdx.x == 0) ws->tr2[par ^ 1] = 0ull;
    }
    if (threadIdx.x < ECO_NLOSS) ms.up[threadIdx.x] = upstream[threadIdx.x];

    // ---- pass 1: per channel sum g, sum x, sum x^2, sum g x, softplus remainder, focal (log2 units) ----------------
    {
        f2 acc[M_NACC];
#pragma unroll
        for (int k = 0; k < M_NACC; ++k) acc[k] = splat(0.f);
        int since_flush = 0;
        int kk = tr.t_lo % tr.tpp;
        const uint32_t my = sbase + threadIdx.x * 8;
        for (int k = 0; k < ntiles; ++k) {
            f2 z[3], g[3];
            consume_tile(my, ms.ps, k, lane, z, g);
            if ((int64_t)kk * kTP + 2 * (int)threadIdx.x < a.HW) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f2* ac = &acc[M_PER * c];
                    const f2 x = PROB ? z[c] : sigmoid_fast2(z[c]);
                    const f2 t = mul2(x, x);
                    ac[M_G] = add2(ac[M_G], g[c]);
                    ac[M_X] = add2(ac[M_X], x);
                    ac[M_XX] = add2(ac[M_XX], t);
                    ac[M_GX] = fma2(g[c], x, ac[M_GX]);
                    f2 q = fma2(t, splat(kSpR2), splat(kSpR1));
                    q = fma2(q, t, splat(kSpR0));
                    ac[M_R] = fma2(mul2(t, t), q, ac[M_R]);
                    const f2 om = fma2(x, splat(-1.0f), splat(1.0f));
                    const f2 sq = make_float2(sqrt_approx(om.x), sqrt_approx(om.y));
                    const f2 be = add2(x, splat(kEps));
                    const f2 lg = make_float2(lg2_approx(be.x), lg2_approx(be.y));
                    ac[M_FL] = fma2(mul2(om, sq), lg, ac[M_FL]);
                }
            }
            if (++kk == tr.tpp) kk = 0;
            if (++since_flush == 2 * kFlushTiles) { mc_flush(acc, ms.warp_slots[warp], lane); since_flush = 0; }   // packed: 64 values per fp32 lane
        }
        mc_flush(acc, ms.warp_slots[warp], lane);
    }
    ECO_TL(1);
    csync();
    if (threadIdx.x < 32) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kCWarps; ++w) v += ms.warp_slots[w][threadIdx.x];
        ms.sums[threadIdx.x] = v;
    }
    csync();
    if (threadIdx.x < kMcSums) {
        const int64_t last = a.HW - (int64_t)(tr.tpp - 1) * kTP;
        const int n_last = tr.t_hi / tr.tpp - tr.t_lo / tr.tpp;
        const double npix = (double)((int64_t)(ntiles - n_last) * kTP + (int64_t)n_last * last);
        double v;
        if (threadIdx.x == 0) v = npix;
        else {
            const int c = (threadIdx.x - 1) / M_PER, k = (threadIdx.x - 1) % M_PER;
            const double* s = ms.sums + M_PER * c;
            if (k == M_R) v = npix * kLn2d + 0.5 * s[M_X] + 0.125 * s[M_XX] + s[M_R];   // sum softplus(x), x in [0,1]
            else if (k == M_FL) v = -kLn2d * s[M_FL];
            else v = s[k];
        }
        fix_add(ws->fix1[par][blockIdx.x % kFixRep] + 2 * threadIdx.x, v);
        __threadfence();
    }
    csync();
    ECO_TL(2);
    if (threadIdx.x == 0) {
        atomicAdd(&ws->tr2[par], 1ull);
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(&ws->tr2[par]) : "memory");
            if (seen < gridDim.x) __nanosleep(32);
        } while (seen < gridDim.x);
    }
    csync();
    ECO_TL(3);
    if (threadIdx.x < kMcSums) ms.acc[threadIdx.x] = fix_get(ws->fix1[par][0] + 2 * threadIdx.x, 2 * kNAcc);
    csync();

    // ---- closed forms (one warp per loss kind, one lane per channel leaf); every row leaves its Jacobian already weighted
    //      by its upstream gradient, three threads add the rows in a fixed order into the gradient coefficients -------------
    if (threadIdx.x < ECO_NLOSS * 32 && lane < 3) {
        const double* s6 = ms.acc + 1 + M_PER * lane;
        double s[ECO_NSTAT], jrow[ECO_NJAC];
        s[S_N] = ms.acc[0]; s[S_A] = s6[M_G]; s[S_B] = s6[M_X]; s[S_AB] = s6[M_GX]; s[S_BB] = s6[M_XX];
        s[S_SP] = s6[M_R]; s[S_FL] = s6[M_FL]; s[S_FLB] = 0.0;
        leaf_closed_form_row(s, 0.0, scale, warp, ms.sl[lane][warp], jrow);
        const float u = warp == 0 ? 0.f : ms.up[warp];
#pragma unroll
        for (int j = 0; j < ECO_NJAC; ++j) ms.jac_s[lane][warp][j] = u != 0.f ? (double)u * jrow[j] : 0.0;   // (an unused loss may have a non-finite Jacobian)
    }
    csync();
    if (threadIdx.x < 3) {
        float c[ECO_NJAC];
#pragma unroll
        for (int j = 0; j < ECO_NJAC; ++j) {
            double v = 0.0;
#pragma unroll
            for (int k = 1; k < ECO_NLOSS; ++k) v += ms.jac_s[threadIdx.x][k][j];
            c[j] = (float)(j == 3 ? 2.0 * v : v);
        }
        // LeafCoef order: sa, sb, sab, 2 sbb, sp, fl, flb
        ms.ua[threadIdx.x] = make_float4(c[1] + 0.5f * c[4], c[2], c[3], c[4]);
        ms.ufl[threadIdx.x] = c[5];
    }
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + ECO_NLOSS)
        losses_out[threadIdx.x - 32] = (float)(ms.sl[0][threadIdx.x - 32] + ms.sl[1][threadIdx.x - 32] + ms.sl[2][threadIdx.x - 32]);
    csync();

    ECO_TL(4);
    // ---- pass 2: d(sum_k upstream_k loss_k)/d logits, walking this CTA's tiles backwards --------------------------------
    const bool need_sig = ms.up[1] != 0.f, need_fl = ms.up[2] != 0.f;
    if (need_fl) {
        if (need_sig) mc_grad_consume<true, true, PROB>(ga, tr, sbase, ms, ntiles);
        else mc_grad_consume<false, true, PROB>(ga, tr, sbase, ms, ntiles);
    } else {
        if (need_sig) mc_grad_consume<true, false, PROB>(ga, tr, sbase, ms, ntiles);
        else mc_grad_consume<false, false, PROB>(ga, tr, sbase, ms, ntiles);
    }
    ECO_TL(5);
    if (blockIdx.x == 0 && threadIdx.x == 0) ws->step = (unsigned int)par + 1u;   // every CTA read `step` before it arrived
}

}  // namespace v2
}  // namespace eco
