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
// One CTA = one 32 x 16 tile of one output frame: input patch -> shared memory (4-byte loads, whatever the alignment of the
// rows), horizontal pass into a byte buffer in shared memory, vertical pass + table look-up, coalesced 128-byte stores.
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

__host__ __device__ inline int frames_pitch(int patch_cols) { return ((patch_cols + kFrPadTaps) * 3 + 3 + 4) / 4 * 4; }
__host__ __device__ inline size_t frames_smem_bytes(int patch_cols, int patch_rows, int ksx, int ksy) {
    size_t b = (size_t)patch_rows * frames_pitch(patch_cols);           // patch
    b += (size_t)patch_rows * (kFrTW * 3);                              // rows after the horizontal pass
    b = (b + 15) / 16 * 16;
    b += (size_t)(kFrTW * ksx + kFrTH * ksy + 2 * kFrTW + 2 * kFrTH) * 4;   // coefficient rows, bounds
    b += 3 * 256 * 4;                                                    // byte -> normalised float
    return b;
}

__device__ __forceinline__ uint32_t clip8(int v) {   // Resample.c: clip8(in) = clip8_lookups[in >> PRECISION_BITS]
    v >>= kFrPrecision;
    return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// the aligned 4-byte word at q; the first / last word of the whole buffer byte by byte: nothing outside the caller's memory
__device__ __forceinline__ uint32_t frames_word(const uint8_t* q, const uint8_t* lo, const uint8_t* hi) {
    if (q >= lo && q + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(q));
    uint32_t v = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (q + b >= lo && q + b < hi) v |= (uint32_t)__ldg(q + b) << (8 * b);
    return v;
}

// KX: taps of the horizontal pass held in registers (>= ksx; 0 = any number, read from shared memory)
template <int KX>
__global__ void __launch_bounds__(kFrThreads)
frames_preprocess_kernel(FrameArgs p) {
    extern __shared__ __align__(16) unsigned char fr_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kFrTW, y0 = blockIdx.y * kFrTH, n = blockIdx.z;
    const int tw = min(kFrTW, p.Wout - x0), th = min(kFrTH, p.Hout - y0);
    unsigned char* patch = fr_smem;
    unsigned char* hbuf = patch + (size_t)p.patch_rows * p.pitch;
    size_t off = ((size_t)p.patch_rows * p.pitch + (size_t)p.patch_rows * (kFrTW * 3) + 15) / 16 * 16;
    int* kxs = reinterpret_cast<int*>(fr_smem + off);
    int* kys = kxs + kFrTW * p.ksx;
    int* xbs = kys + kFrTH * p.ksy;
    int* ybs = xbs + 2 * kFrTW;
    float* lut = reinterpret_cast<float*>(ybs + 2 * kFrTH);
    __shared__ int ext[4];   // cx0, ncols, ry0, nrows

    for (int i = tid; i < kFrTW * p.ksx; i += kFrThreads) kxs[i] = i < tw * p.ksx ? p.kx[(size_t)x0 * p.ksx + i] : 0;
    for (int i = tid; i < th * p.ksy; i += kFrThreads) kys[i] = p.ky[(size_t)y0 * p.ksy + i];
    for (int i = tid; i < 2 * kFrTW; i += kFrThreads) xbs[i] = i < 2 * tw ? p.xb[2 * x0 + i] : 0;
    for (int i = tid; i < 2 * th; i += kFrThreads) ybs[i] = p.yb[2 * y0 + i];
    for (int i = tid; i < 3 * 256; i += kFrThreads) lut[i] = p.lut[i];
    __syncthreads();
    if (tid == 0) {
        int lo = xbs[0], hi = xbs[0] + xbs[1];
        for (int x = 1; x < tw; ++x) { lo = min(lo, xbs[2 * x]); hi = max(hi, xbs[2 * x] + xbs[2 * x + 1]); }
        ext[0] = lo; ext[1] = min(hi - lo, p.patch_cols);   // (eco_frames_plan sized the patch for every tile)
        lo = ybs[0]; hi = ybs[0] + ybs[1];
        for (int y = 1; y < th; ++y) { lo = min(lo, ybs[2 * y]); hi = max(hi, ybs[2 * y] + ybs[2 * y + 1]); }
        ext[2] = lo; ext[3] = min(hi - lo, p.patch_rows);
    }
    __syncthreads();
    const int cx0 = ext[0], ncols = ext[1], ry0 = ext[2], nrows = ext[3];

    // ---- input patch -> shared memory: one warp per row, aligned 4-byte loads shifted so that every row starts at byte 0 --
    const uint8_t* fbase = p.src + (int64_t)n * p.frame_stride;
    for (int r = warp; r < nrows; r += kFrThreads / 32) {
        const uint8_t* g0 = fbase + (int64_t)(ry0 + r) * p.row_stride + (int64_t)cx0 * 3;
        const int o = (int)(reinterpret_cast<uintptr_t>(g0) & 3);
        const uint8_t* wbase = g0 - o;
        const int nwords = (ncols * 3 + 3) >> 2;
        uint32_t* dst = reinterpret_cast<uint32_t*>(patch + (size_t)r * p.pitch);
        // (every word of the row and the one after it inside the caller's buffer: plain loads, the neighbour by shuffle)
        const bool inside = wbase >= p.src && wbase + 4 * (nwords + 1) <= p.src_end;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const int w = w0 + lane;
            uint32_t a = 0u;
            if (w <= nwords) a = inside ? __ldg(reinterpret_cast<const uint32_t*>(wbase) + w) : frames_word(wbase + 4 * w, p.src, p.src_end);
            uint32_t b = __shfl_down_sync(0xffffffffu, a, 1);
            if (lane == 31 && o && w < nwords) b = inside ? __ldg(reinterpret_cast<const uint32_t*>(wbase) + w + 1) : frames_word(wbase + 4 * w + 4, p.src, p.src_end);
            if (w < nwords) dst[w] = __funnelshift_r(a, b, 8 * o);
        }
    }
    __syncthreads();

    // ---- horizontal pass: (patch row, output column, channel) -> byte.  A thread keeps ONE (column, channel) and its
    //      coefficients in registers and walks down the rows of its half of the patch --------------------------------------
    {
        const int xc = tid % (kFrTW * 3), half = tid / (kFrTW * 3);
        const int x = xc / 3, c = xc - 3 * x;
        if (x < tw) {
            const int first = xbs[2 * x] - cx0;
            const unsigned char* pcol = patch + first * 3 + c;
            if (KX > 0) {
                int k[KX > 0 ? KX : 1];
#pragma unroll
                for (int j = 0; j < KX; ++j) k[j] = j < p.ksx ? kxs[x * p.ksx + j] : 0;   // (zero beyond the tap count)
                for (int r = half; r < nrows; r += 2) {
                    const unsigned char* prow = pcol + (size_t)r * p.pitch;
                    int acc = 1 << (kFrPrecision - 1);
#pragma unroll
                    for (int j = 0; j < KX; ++j) acc += k[j] * (int)prow[3 * j];
                    hbuf[r * (kFrTW * 3) + xc] = (unsigned char)clip8(acc);
                }
            } else {
                const int cnt = xbs[2 * x + 1];
                const int* k = kxs + x * p.ksx;
                for (int r = half; r < nrows; r += 2) {
                    const unsigned char* prow = pcol + (size_t)r * p.pitch;
                    int acc = 1 << (kFrPrecision - 1);
                    for (int j = 0; j < cnt; ++j) acc += k[j] * (int)prow[3 * j];
                    hbuf[r * (kFrTW * 3) + xc] = (unsigned char)clip8(acc);
                }
            }
        }
    }
    __syncthreads();

    // ---- vertical pass + byte -> normalised float; a warp writes 32 consecutive floats of one (channel, row) -----------
    const int x = lane;
    if (x < tw) {
        for (int yc = warp; yc < th * 3; yc += kFrThreads / 32) {
            const int y = yc / 3, c = yc - 3 * y;
            const int first = ybs[2 * y] - ry0, cnt = ybs[2 * y + 1];
            const unsigned char* col = hbuf + (size_t)first * (kFrTW * 3) + x * 3 + c;
            const int* k = kys + y * p.ksy;
            int acc = 1 << (kFrPrecision - 1);
            for (int j = 0; j < cnt; ++j) acc += k[j] * (int)col[(size_t)j * (kFrTW * 3)];
            p.out[(((int64_t)n * 3 + c) * p.Hout + (y0 + y)) * p.Wout + (x0 + x)] = lut[c * 256 + clip8(acc)];
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
    void (*kernel)(FrameArgs) = ksx <= 3 ? frames_preprocess_kernel<3> : ksx <= 5 ? frames_preprocess_kernel<5>
                              : ksx <= 7 ? frames_preprocess_kernel<7> : ksx <= 9 ? frames_preprocess_kernel<9>
                              : ksx <= 13 ? frames_preprocess_kernel<13> : ksx <= 17 ? frames_preprocess_kernel<17>
                              : frames_preprocess_kernel<0>;
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
