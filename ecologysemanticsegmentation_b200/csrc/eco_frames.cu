// Frame pre-processing in front of the network:
//   ess/test_video.py:70-78   transforms.Resize((256, 256)) -> transforms.ToTensor() -> transforms.Normalize(mean, std)
// on a PIL RGB image.  The reference resizes on the host (Pillow's two 8-bit passes), converts to float, normalises and
// only then copies 4 B/element to the GPU; here the uint8 frames (3 B/pixel) go to the device as they are and ONE kernel
// does both resampling passes, the /255 and the normalisation: it reads every input byte once (plus the tile halos) and
// writes the fp32 CHW tensor the network reads.  HBM-bound in principle: 3 B per input pixel + 12 B per output pixel.
//
// The arithmetic is Pillow's, bit for bit (third-party code, not part of the reference tree; Pillow 12.2.0
// src/libImaging/Resample.c): a separable triangle filter whose support grows with the down-scaling factor
// (precompute_coeffs), coefficients normalised in double and rounded to 22-bit fixed point (normalize_coeffs_8bpc),
// horizontal pass first with one rounding to uint8 (ImagingResampleHorizontal_8bpc), then the vertical pass with another
// (ImagingResampleVertical_8bpc).  The tables are built on the host by eco_frames_plan (plain C double arithmetic in
// Pillow's operation order) and the per-channel map byte -> (byte / 255 - mean) / std is a 3 x 256 table of IEEE
// float32 operations (torchvision F.to_tensor / F.normalize on the CPU: true division, no reciprocal).
//
// One CTA = one 32 x 16 tile of one output frame: input patch -> shared memory by ONE 1-D TMA bulk copy per patch row
// (cp.async.bulk from the 16-byte aligned superset of the row; the row then starts `o` bytes into its shared-memory line),
// horizontal pass into a byte buffer in shared memory, vertical pass (taps unrolled, one coefficient fetch per output row for
// its three channels) + table look-up, coalesced 128-byte stores.  Rows whose aligned superset would leave the caller's
// buffer (first / last rows of the whole batch) are fetched word by word into the same layout.
#include <cmath>

#include "eco_common.cuh"

namespace eco {

constexpr int kFrTW = 32, kFrTH = 16, kFrThreads = 192;   // 192 = 2 x (32 columns x 3 channels) for the horizontal pass
constexpr int kFrPrecision = 32 - 8 - 2;   // Resample.c: PRECISION_BITS
constexpr int kFrPadTaps = 17;             // the horizontal pass may read (and multiply by 0) this many pixels past a row

struct FrameArgs {
    const uint8_t* src;
    const uint8_t* src_end;          // one past the last byte of the frames
    int64_t frame_stride, row_stride;   // bytes
    int32_t N, Hin, Win, Hout, Wout, ksx, ksy;
    const int32_t *xb, *kx, *yb, *ky;
    const float* lut;
    float* out;
    int32_t patch_cols, patch_rows, pitch;   // shared-memory patch: rows x pitch bytes (pitch % 4 == 0)
};

// a patch row in shared memory: up to 15 bytes in front of the row (16-byte aligned source), the row, the horizontal pass's
// read-ahead, rounded up to the 16-byte granule of the bulk copy
__host__ __device__ inline int frames_pitch(int patch_cols) { return (15 + (patch_cols + kFrPadTaps) * 3 + 15) / 16 * 16; }
__host__ __device__ inline size_t frames_smem_bytes(int patch_cols, int patch_rows, int ksx, int ksy) {
    size_t b = (size_t)patch_rows * frames_pitch(patch_cols);           // patch
    b += (size_t)(patch_rows + kFrPadTaps) * (kFrTW * 3);               // rows after the horizontal pass (+ the vertical pass's zero-weighted read-ahead)
    b = (b + 15) / 16 * 16;
    b += (size_t)(kFrTW * ksx + kFrTH * ksy + 2 * kFrTW + 2 * kFrTH) * 4;   // coefficient rows, bounds
    return b;
}

__device__ __forceinline__ uint32_t clip8(int v) {   // Resample.c: clip8(in) = clip8_lookups[in >> PRECISION_BITS]
    v >>= kFrPrecision;
    return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// the aligned 4-byte word at q; the first / last word of the whole buffer byte by byte: nothing outside the caller's memory
__device__ __noinline__ uint32_t frames_word(const uint8_t* q, const uint8_t* lo, const uint8_t* hi) {
    if (q >= lo && q + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(q));
    uint32_t v = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (q + b >= lo && q + b < hi) v |= (uint32_t)__ldg(q + b) << (8 * b);
    return v;
}

__device__ __forceinline__ uint32_t fr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// KX / KY: taps of the horizontal / vertical pass held in registers (>= ksx / ksy; 0 = any number, read from shared memory)
template <int KX, int KY>
__global__ void __launch_bounds__(kFrThreads)
frames_preprocess_kernel(FrameArgs p) {
    extern __shared__ __align__(16) unsigned char fr_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kFrTW, y0 = blockIdx.y * kFrTH, n = blockIdx.z;
    const int tw = min(kFrTW, p.Wout - x0), th = min(kFrTH, p.Hout - y0);
    unsigned char* patch = fr_smem;
    unsigned char* hbuf = patch + (size_t)p.patch_rows * p.pitch;
    size_t off = ((size_t)p.patch_rows * p.pitch + (size_t)(p.patch_rows + kFrPadTaps) * (kFrTW * 3) + 15) / 16 * 16;
    int* kxs = reinterpret_cast<int*>(fr_smem + off);
    int* kys = kxs + kFrTW * p.ksx;
    int* xbs = kys + kFrTH * p.ksy;
    int* ybs = xbs + 2 * kFrTW;
    const float* __restrict__ lut = p.lut;   // 3 KB, read through L1 (a copy per CTA cost more than its 1536 look-ups)
    __shared__ int ext[4];   // cx0, ncols, ry0, nrows

    // Tap counts held in registers (the usual case): every thread fetches its own coefficients and bounds straight from the
    // tables in global memory (L1 / L2 hits, under the patch copies); only the any-number-of-taps instantiations stage the
    // tile's table rows in shared memory
    if (KX == 0) {
        for (int i = tid; i < kFrTW * p.ksx; i += kFrThreads) kxs[i] = i < tw * p.ksx ? p.kx[(size_t)x0 * p.ksx + i] : 0;
        for (int i = tid; i < 2 * kFrTW; i += kFrThreads) xbs[i] = i < 2 * tw ? p.xb[2 * x0 + i] : 0;
    }
    if (KY == 0) {
        for (int i = tid; i < th * p.ksy; i += kFrThreads) kys[i] = p.ky[(size_t)y0 * p.ksy + i];
        for (int i = tid; i < 2 * th; i += kFrThreads) ybs[i] = p.yb[2 * y0 + i];
    }
    if (warp == 0) {   // extent of the input patch: min / max over the tile's columns and rows (kFrTW == 32 lanes)
        const bool hx = lane < tw, hy = lane < th;
        const int xb0 = hx ? __ldg(p.xb + 2 * (x0 + lane)) : 0, xb1 = hx ? __ldg(p.xb + 2 * (x0 + lane) + 1) : 0;
        const int yb0 = hy ? __ldg(p.yb + 2 * (y0 + lane)) : 0, yb1 = hy ? __ldg(p.yb + 2 * (y0 + lane) + 1) : 0;
        const int xlo = __reduce_min_sync(0xffffffffu, hx ? xb0 : 0x7fffffff);
        const int xhi = __reduce_max_sync(0xffffffffu, hx ? xb0 + xb1 : -0x7fffffff);
        const int ylo = __reduce_min_sync(0xffffffffu, hy ? yb0 : 0x7fffffff);
        const int yhi = __reduce_max_sync(0xffffffffu, hy ? yb0 + yb1 : -0x7fffffff);
        if (lane == 0) {
            ext[0] = xlo; ext[1] = min(xhi - xlo, p.patch_cols);   // (eco_frames_plan sized the patch for every tile)
            ext[2] = ylo; ext[3] = min(yhi - ylo, p.patch_rows);
        }
    }
    __syncthreads();
    const int cx0 = ext[0], ncols = ext[1], ry0 = ext[2], nrows = ext[3];

    // ---- input patch -> shared memory.  Row r of the patch starts at fbase + (ry0 + r) * row_stride + cx0 * 3; its 16-byte
    //      aligned superset goes to patch + r * pitch, so the row itself starts row_off(r) bytes into that line ------------------
    const uint8_t* fbase = p.src + (int64_t)n * p.frame_stride;
    const uint8_t* g00 = fbase + (int64_t)ry0 * p.row_stride + (int64_t)cx0 * 3;
    const uint32_t a00 = (uint32_t)(reinterpret_cast<uintptr_t>(g00) & 15u), rs15 = (uint32_t)(p.row_stride & 15);
    auto row_off = [&](int r) { return (int)((a00 + (uint32_t)r * rs15) & 15u); };
    __shared__ __align__(8) unsigned long long tma_bar;
    const int row_bytes = ncols * 3;
    auto row_fast = [&](int r) {   // the row's 16-byte aligned superset lies inside the caller's buffer
        const int o = row_off(r);
        const uint8_t* src16 = g00 + (int64_t)r * p.row_stride - o;
        return src16 >= p.src && src16 + ((o + row_bytes + 15) & ~15) <= p.src_end;
    };
    // Rows lie row_stride >= row_bytes apart, so when the patch's first and last row are inside by this test and a row holds
    // at least the 15 bytes a superset can stick out, every row in between is inside too: one CTA-uniform test instead of one
    // per row (only tiles on the first / last rows of the whole batch fail it)
    const bool all_fast = nrows > 0 && row_bytes >= 16 && row_fast(0) && row_fast(nrows - 1);
    if (warp == 0) {   // the lanes of warp 0 issue the rows' copies side by side
        const uint32_t bar = fr_smem_u32(&tma_bar);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        uint32_t mine = 0;
        for (int r = lane; r < nrows; r += 32)
            if (all_fast || row_fast(r)) mine += (uint32_t)((row_off(r) + row_bytes + 15) & ~15);
        const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
        __syncwarp();
        for (int r = lane; r < nrows; r += 32) {
            if (!all_fast && !row_fast(r)) continue;
            const int o = row_off(r);
            const uint8_t* src16 = g00 + (int64_t)r * p.row_stride - o;
            const uint32_t bytes = (uint32_t)((o + row_bytes + 15) & ~15);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             fr_smem_u32(patch + (size_t)r * p.pitch)), "l"(src16), "r"(bytes), "r"(bar)
                         : "memory");
        }
    }
    if (!all_fast) {
        // (rare) rows at the very start / end of the caller's buffer: word by word, nothing outside the caller's memory is read
        for (int r = warp; r < nrows; r += kFrThreads / 32) {
            if (row_fast(r)) continue;
            const int o = row_off(r);
            const uint8_t* src16 = g00 + (int64_t)r * p.row_stride - o;
            const int bytes = (o + row_bytes + 15) & ~15;
            uint32_t* dst = reinterpret_cast<uint32_t*>(patch + (size_t)r * p.pitch);
            for (int w = lane; w < bytes / 4; w += 32) dst[w] = frames_word(src16 + 4 * w, p.src, p.src_end);
        }
    }
    __syncthreads();   // (the barrier's initialisation is visible to every thread; the word-by-word rows are in place)
    {
        const uint32_t bar = fr_smem_u32(&tma_bar);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(ok) : "r"(bar) : "memory");
        }
    }

    // ---- horizontal pass: (patch row, output column, channel) -> byte.  A thread keeps ONE (column, channel) and its
    //      coefficients in registers and walks down the rows of its half of the patch --------------------------------------
    {
        const int xc = tid % (kFrTW * 3), half = tid / (kFrTW * 3);
        const int x = xc / 3, c = xc - 3 * x;
        if (x < tw) {
            const int first = (KX > 0 ? __ldg(p.xb + 2 * (x0 + x)) : xbs[2 * x]) - cx0;
            const unsigned char* pcol = patch + first * 3 + c;
            if (KX > 0) {
                int k[KX > 0 ? KX : 1];
                const int32_t* kg = p.kx + (size_t)(x0 + x) * p.ksx;
#pragma unroll
                for (int j = 0; j < KX; ++j) k[j] = j < p.ksx ? __ldg(kg + j) : 0;   // (zero beyond the tap count)
                // (row pointer, row offset and output pointer advance incrementally: two rows per step)
                const unsigned char* pline = pcol + half * p.pitch;
                unsigned char* hout = hbuf + half * (kFrTW * 3) + xc;
                uint32_t o = a00 + (uint32_t)half * rs15;
                const uint32_t o_step = 2u * rs15;
                const int line_step = 2 * p.pitch;
                for (int r = half; r < nrows; r += 2) {
                    const unsigned char* prow = pline + (o & 15u);
                    int acc = 1 << (kFrPrecision - 1);
#pragma unroll
                    for (int j = 0; j < KX; ++j) acc += k[j] * (int)prow[3 * j];
                    *hout = (unsigned char)clip8(acc);
                    pline += line_step; hout += 2 * (kFrTW * 3); o += o_step;
                }
            } else {
                const int cnt = xbs[2 * x + 1];
                const int* k = kxs + x * p.ksx;
                for (int r = half; r < nrows; r += 2) {
                    const unsigned char* prow = pcol + r * p.pitch + row_off(r);
                    int acc = 1 << (kFrPrecision - 1);
                    for (int j = 0; j < cnt; ++j) acc += k[j] * (int)prow[3 * j];
                    hbuf[r * (kFrTW * 3) + xc] = (unsigned char)clip8(acc);
                }
            }
        }
    }
    __syncthreads();

    // ---- vertical pass + byte -> normalised float; a warp takes whole output rows: the row's coefficients are fetched once
    //      for its three channels, and the warp writes 32 consecutive floats of one (channel, row) at a time -----------------
    const int x = lane;
    if (x < tw) {
        for (int y = warp; y < th; y += kFrThreads / 32) {
            const int first = (KY > 0 ? __ldg(p.yb + 2 * (y0 + y)) : ybs[2 * y]) - ry0;
            const int cnt = KY > 0 ? 0 : ybs[2 * y + 1];
            const unsigned char* col0 = hbuf + first * (kFrTW * 3) + x * 3;
            const int* k = KY > 0 ? p.ky + (size_t)(y0 + y) * p.ksy : kys + y * p.ksy;
            float* orow = p.out + (((int64_t)n * 3) * p.Hout + (y0 + y)) * p.Wout + (x0 + x);
            if (KY > 0) {
                int kk[KY > 0 ? KY : 1];
#pragma unroll
                for (int j = 0; j < KY; ++j) kk[j] = j < p.ksy ? __ldg(k + j) : 0;   // (the table rows are zero beyond the tap count)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int acc = 1 << (kFrPrecision - 1);
#pragma unroll
                    for (int j = 0; j < KY; ++j) acc += kk[j] * (int)col0[j * (kFrTW * 3) + c];
                    orow[(int64_t)c * p.Hout * p.Wout] = __ldg(lut + c * 256 + clip8(acc));
                }
            } else {
                for (int c = 0; c < 3; ++c) {
                    int acc = 1 << (kFrPrecision - 1);
                    for (int j = 0; j < cnt; ++j) acc += k[j] * (int)col0[j * (kFrTW * 3) + c];
                    orow[(int64_t)c * p.Hout * p.Wout] = __ldg(lut + c * 256 + clip8(acc));
                }
            }
        }
    }
}

// Resample.c: precompute_coeffs (bilinear: support 1, filter 1 - |x| on (-1, 1)) + normalize_coeffs_8bpc for one axis
static int frames_axis(int in_size, int out_size, int ksize, int32_t* bounds, int32_t* kk) {
    const double scale = (double)in_size / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;
    const double ss = 1.0 / filterscale;
    double w[1024];
    if (ksize > 1024) return -1;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double v = a < 1.0 ? 1.0 - a : 0.0;
            w[x] = v;
            ww += v;
        }
        int32_t* k = kk + (size_t)xx * ksize;
        for (int x = 0; x < ksize; ++x) k[x] = 0;
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            k[x] = v < 0 ? (int)(-0.5 + v * (1 << kFrPrecision)) : (int)(0.5 + v * (1 << kFrPrecision));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return 0;
}
static int frames_ksize(int in_size, int out_size) {
    double filterscale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(1.0 * filterscale) * 2 + 1;
}
static int frames_patch_extent(const int32_t* bounds, int out_size, int tile) {
    int best = 0;
    for (int t0 = 0; t0 < out_size; t0 += tile) {
        int lo = bounds[2 * t0], hi = lo + bounds[2 * t0 + 1];
        for (int x = t0; x < out_size && x < t0 + tile; ++x) {
            if (bounds[2 * x] < lo) lo = bounds[2 * x];
            if (bounds[2 * x] + bounds[2 * x + 1] > hi) hi = bounds[2 * x] + bounds[2 * x + 1];
        }
        if (hi - lo > best) best = hi - lo;
    }
    return best;
}

}  // namespace eco

using namespace eco;

extern "C" int eco_frames_plan_sizes(int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, int32_t* ksx, int32_t* ksy) {
    if (Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || !ksx || !ksy) { set_error("bad frame sizes"); return -1; }
    *ksx = frames_ksize(Win, Wout);
    *ksy = frames_ksize(Hin, Hout);
    return 0;
}

extern "C" int eco_frames_plan(int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, const float* mean, const float* stdev,
                               int32_t* xbounds, int32_t* kx, int32_t* ybounds, int32_t* ky, float* lut,
                               int32_t* patch_cols, int32_t* patch_rows) {
    if (Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) { set_error("bad frame sizes"); return -1; }
    if (!mean || !stdev || !xbounds || !kx || !ybounds || !ky || !lut || !patch_cols || !patch_rows) { set_error("null plan output"); return -1; }
    if (frames_axis(Win, Wout, frames_ksize(Win, Wout), xbounds, kx) || frames_axis(Hin, Hout, frames_ksize(Hin, Hout), ybounds, ky)) {
        set_error("down-scaling factor too large (more than 511 taps)");
        return -3;
    }
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            // F.to_tensor: float32(v) / 255; F.normalize: (x - mean) / std -- three IEEE float32 operations
            volatile float a = (float)v / 255.0f;
            volatile float b = a - mean[c];
            lut[c * 256 + v] = b / stdev[c];
        }
    *patch_cols = frames_patch_extent(xbounds, Wout, kFrTW);
    *patch_rows = frames_patch_extent(ybounds, Hout, kFrTH);
    return 0;
}

extern "C" int eco_frames_preprocess(const uint8_t* frames, int32_t N, int32_t Hin, int32_t Win, int64_t frame_stride_bytes,
                                     int64_t row_stride_bytes, const int32_t* xbounds_dev, const int32_t* kx_dev, int32_t ksx,
                                     const int32_t* ybounds_dev, const int32_t* ky_dev, int32_t ksy, int32_t Hout, int32_t Wout,
                                     int32_t patch_cols, int32_t patch_rows, const float* lut_dev, float* out, int device,
                                     void* stream) {
    if (!frames || !out || !xbounds_dev || !kx_dev || !ybounds_dev || !ky_dev || !lut_dev) { set_error("null frames / tables / output"); return -1; }
    if (N <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) { set_error("empty input (N=%d %dx%d -> %dx%d)", N, Hin, Win, Hout, Wout); return -2; }
    if (ksx != frames_ksize(Win, Wout) || ksy != frames_ksize(Hin, Hout) || patch_cols <= 0 || patch_rows <= 0) {
        set_error("tables do not belong to these sizes (build them with eco_frames_plan)");
        return -3;
    }
    if (row_stride_bytes < (int64_t)Win * 3 || frame_stride_bytes < (int64_t)(Hin - 1) * row_stride_bytes + (int64_t)Win * 3) {
        set_error("frame / row strides smaller than the data");
        return -3;
    }
    if (N > 65535) { set_error("at most 65535 frames per call"); return -3; }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cannot select device %d", device); return -6; }
    const size_t smem = frames_smem_bytes(patch_cols, patch_rows, ksx, ksy);
    if (smem > 200 * 1024) { set_error("down-scaling factor too large for one tile's input patch (%zu bytes of shared memory)", smem); return -8; }
    // taps in registers: 3 (up-scaling and no scaling), 5, 7, 9, 13, 17; anything longer reads its taps from shared memory
    auto bucket = [](int k) { return k <= 3 ? 0 : k <= 5 ? 1 : k <= 7 ? 2 : k <= 9 ? 3 : k <= 13 ? 4 : k <= 17 ? 5 : 6; };
    static void (*const kernels[7][7])(FrameArgs) = {
#define ECO_FR_ROW(KX) {frames_preprocess_kernel<KX, 3>, frames_preprocess_kernel<KX, 5>, frames_preprocess_kernel<KX, 7>, \
                        frames_preprocess_kernel<KX, 9>, frames_preprocess_kernel<KX, 13>, frames_preprocess_kernel<KX, 17>, \
                        frames_preprocess_kernel<KX, 0>}
        ECO_FR_ROW(3), ECO_FR_ROW(5), ECO_FR_ROW(7), ECO_FR_ROW(9), ECO_FR_ROW(13), ECO_FR_ROW(17), ECO_FR_ROW(0)
#undef ECO_FR_ROW
    };
    void (*kernel)(FrameArgs) = kernels[bucket(ksx)][bucket(ksy)];
    if (smem > 48 * 1024) {
        int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(smem, frames)");
        if (rc) return rc;
    }
    FrameArgs p{};
    p.src = frames;
    p.src_end = frames + (int64_t)(N - 1) * frame_stride_bytes + (int64_t)(Hin - 1) * row_stride_bytes + (int64_t)Win * 3;
    p.frame_stride = frame_stride_bytes; p.row_stride = row_stride_bytes;
    p.N = N; p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout; p.ksx = ksx; p.ksy = ksy;
    p.xb = xbounds_dev; p.kx = kx_dev; p.yb = ybounds_dev; p.ky = ky_dev; p.lut = lut_dev; p.out = out;
    p.patch_cols = patch_cols; p.patch_rows = patch_rows; p.pitch = frames_pitch(patch_cols);
    dim3 grid((unsigned)((Wout + kFrTW - 1) / kFrTW), (unsigned)((Hout + kFrTH - 1) / kFrTH), (unsigned)N);
    kernel<<<grid, kFrThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_cuda(cudaGetLastError(), "frames_preprocess_kernel launch");
}
