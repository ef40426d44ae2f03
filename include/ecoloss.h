/*
 * ecoloss.h -- C ABI of the B200-native per-pixel loss / Dice-scoring path.
 *
 * The reference (hansk0812/EcologySemanticSegmentation) has no FFI layer: its seam is plain
 * Python.  Each entry point below names the reference code it replaces
 * (paths relative to the reference root, `ess/` = `ecology_semantic_segmentation/`); the
 * Python side of this repo (`ecologysemanticsegmentation_b200/`) binds them with ctypes and
 * re-exposes the reference's own function names and signatures.  INTEGRATION.md shows the stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the library allocates nothing: the caller owns all buffers, including the workspace;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*) of device
 *     `device`; entry points are re-entrant (PyTorch calls backward from its autograd thread);
 *   - return value 0 = success, >0 = cudaError_t, <0 = argument error; `eco_last_error()`
 *     returns a thread-local description;
 *   - tensors are NCHW; a "plane" is the H*W contiguous elements of one (n, c).  A tensor is
 *     described by its base pointer and two element strides: `sn` between images and `sc`
 *     between channels, so `x[:, c:c+1]` slices of a contiguous tensor need no copy;
 *   - dtype codes: ECO_F32 = 0, ECO_BF16 = 1 (logits / probabilities / labels / gradients), ECO_U8 = 2 (byte
 *     masks, accepted for the labels of eco_dice_counts and of eco_composite3_step);
 *   - "slots": `a` is the reference's FIRST positional argument ("gt"), `b` the SECOND ("pred").
 *     Which one is really the label depends on the caller.
 *
 * Statistics vector of one (a, b) leaf (float64[ECO_NSTAT]), additive across shards:
 *   [0] n  [1] sum a  [2] sum b  [3] sum a*b  [4] sum b*b
 *   [5] sum max(b,0)+log(1+exp(-|b|))  [6] sum -(1-b)^1.5 log(b+1e-7)  [7] sum -b^1.5 log(1-b+1e-7)
 * Loss order everywhere (ess/loss_composite.py:39):
 *   [ce, bce, focal, dice, generalized_dice, twersky, focal_dice]
 */
#ifndef ECOLOSS_H
#define ECOLOSS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECO_F32 0
#define ECO_BF16 1
#define ECO_U8 2 /* labels only (eco_dice_counts, eco_composite3_step): masks stored as bytes */

#define ECO_NSTAT 8
#define ECO_NLOSS 7
#define ECO_NJAC 7 /* d loss_k / d stat[1..7] */

/* flags for eco_pair_* */
#define ECO_A_LOGIT 1      /* slot a holds logits: sigmoid applied on load, gradient chained to the logit */
#define ECO_B_LOGIT 2      /* same for slot b */
#define ECO_NEED_BG 4      /* background_weight != 0: also accumulate stat [7] */
#define ECO_UNIT_RANGE_B 8 /* caller guarantees b in [0,1] (e.g. b = sigmoid(.)): lets max(b,0) collapse to b */

/* A strided view of C planes-per-image over N images. */
typedef struct EcoView {
    const void* ptr;
    int64_t sn; /* element stride between images   */
    int64_t sc; /* element stride between channels */
    int32_t dtype;
    int32_t _pad;
} EcoView;

/* Mutable flavour for gradient outputs (ptr may be NULL = not requested). */
typedef struct EcoOut {
    void* ptr;
    int64_t sn;
    int64_t sc;
    int32_t dtype;
    int32_t _pad;
} EcoOut;

const char* eco_version(void);
const char* eco_last_error(void);
/* Number of SMs of `device` (grid sizing is a multiple of it); <0 on error. */
int eco_sm_count(int device);

/* ------------------------------------------------------------------------------------------
 * Pair-leaf engine: C independent (a_c, b_c) leaves in one pass.
 * Replaces, per leaf, the 7 calls of ess/loss_composite.py:32-39 (= ess/train_multiclass.py:269-272):
 *   cross_entropy_loss(bce=True) (ess/loss_functions.py:26,34 + ess/__init__.py:24), cross_entropy_loss
 *   (bce=False, :29-30, identically 0 on one channel), focal_loss (:46), classification_dice_loss (:110)
 *   -> dice_loss (:52), twersky_loss (:82), focal_dice_coefficient (:96);
 * and, used one statistic at a time, each of those primitives when called on its own.
 * ------------------------------------------------------------------------------------------ */

/* Workspace bytes needed by eco_pair_stats for C channels (partials + arrival counters). */
int64_t eco_pair_ws_bytes(int32_t C);

/* Pass 1.  sums_out: float64[C][ECO_NSTAT] (this shard's sums; n = N*HW).  Deterministic. */
int eco_pair_stats(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                   void* ws, int64_t ws_bytes, double* sums_out, int device, void* stream);

/* sums -> losses and Jacobian.  `sums` may have been all-reduced across shards first (its [0] is the
 * global n).  scale_host[c] multiplies leaf c's 7 losses (2.0 for ess/loss_composite.py:40's doubling).
 * losses_out: float32[C][7]; total_out: float32[7] = sum over c (may be NULL);
 * jac_out: float64[C][7][ECO_NJAC] = d(scale*loss_k)/d stat[1+s]. */
int eco_pair_finalize(const double* sums, int32_t C, double background_weight, const double* scale_host,
                      float* losses_out, float* total_out, double* jac_out, int device, void* stream);

/* Pass 2.  upstream: float32[7] device (dT/dloss_k, shared by all channels; ce's entry is ignored).
 * ga/gb: gradient w.r.t. slot a / slot b (w.r.t. the logit where the slot is flagged as logit);
 * either may have ptr == NULL.  accumulate != 0 adds into the outputs instead of overwriting. */
int eco_pair_grad(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                  const double* jac, const float* upstream, const EcoOut* ga, const EcoOut* gb, int32_t accumulate,
                  int device, void* stream);

/* The same three calls with the shape parameters the reference's primitives accept as keyword arguments:
 * focal_loss(gamma=1.5) (ess/loss_functions.py:46), twersky_loss(alpha=0.5, beta=0.3) (:82),
 * focal_dice_coefficient(gamma=1.8) (:96).  shape_host == NULL = those defaults (= the calls above).  A non-default
 * focal_gamma changes stat [6]/[7] to sum -(1-b)^gamma log(b+1e-7) / sum -b^gamma log(1-b+1e-7), so the SAME shape
 * must be passed to stats, finalize and grad of one evaluation. */
typedef struct EcoLeafShape {
    double focal_gamma;
    double tversky_alpha;
    double tversky_beta;
    double focal_dice_gamma;
} EcoLeafShape;

int eco_pair_stats_shaped(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                          const EcoLeafShape* shape_host, void* ws, int64_t ws_bytes, double* sums_out, int device,
                          void* stream);
int eco_pair_finalize_shaped(const double* sums, int32_t C, double background_weight, const double* scale_host,
                             const EcoLeafShape* shape_host, float* losses_out, float* total_out, double* jac_out,
                             int device, void* stream);
int eco_pair_grad_shaped(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                         const EcoLeafShape* shape_host, const double* jac, const float* upstream, const EcoOut* ga,
                         const EcoOut* gb, int32_t accumulate, int device, void* stream);

/* One step of C independent leaves in ONE cooperative launch (C <= 64, grid = one resident wave): pass 1 ->
 * per-channel hand-over of the sums (closed forms and gradient coefficients by the channel's last CTA) -> pass 2 over
 * the same tiles.  This is `loss.backward()` of ess/train_multiclass.py:139-147 for the single-organ configuration
 * (ORGANS=whole_body: C == 1, a = prediction, b = label, background_weight honoured, ess/loss_composite.py:32-40) and
 * for any plain multi-channel losses_fn (a = labels, b = predictions, :28), with upstream = dT/dloss_k known up front.
 * scale multiplies every leaf's 7 losses (2.0 for ess/loss_composite.py:40).  sums_out: float64[C][ECO_NSTAT];
 * losses_out: float32[7] totals over the channels; ga / gb as in eco_pair_grad (overwritten, not accumulated).
 * ws: eco_pair_fused_ws_bytes(C) bytes, zeroed once.  Returns -8 when C leaves do not fit one resident wave. */
int64_t eco_pair_fused_ws_bytes(int32_t C);
int eco_pair_fused(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                   double background_weight, double scale, const EcoLeafShape* shape_host, const float* upstream,
                   void* ws, int64_t ws_bytes, double* sums_out, float* losses_out, const EcoOut* ga, const EcoOut* gb,
                   int device, void* stream);
/* The same with the "only if changed" form of the drop-in autograd path: upstream_prev (device float32[7] or NULL) are the
 * weights the output buffers already hold the step for; the kernel compares the two vectors on the device and returns at
 * once when they are equal (the backward of a step whose forward anticipated the weights of ess/train_multiclass.py:145). */
int eco_pair_fused_ex(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, uint32_t flags,
                      double background_weight, double scale, const EcoLeafShape* shape_host, const float* upstream,
                      const float* upstream_prev, void* ws, int64_t ws_bytes, double* sums_out, float* losses_out,
                      const EcoOut* ga, const EcoOut* gb, int device, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3-organ composite loss: ess/loss_composite.py:21-94 `losses_fn(x, g, composite_set_theory=True)`
 * with C == 3 (whole_body, ventral+dorsal, dorsal): the 3 per-channel leaves (:28) plus, per organ
 * pair i<j, the six intersection_loss / union_loss leaves (:56-81, :87-94) = 21 leaves, fused.
 * x: probabilities, or logits when from_logits != 0 (then sigmoid of ess/train_multiclass.py:134 is
 * fused and the gradient is w.r.t. the logits).  g: labels.  Gradient is produced for x only.
 * ------------------------------------------------------------------------------------------ */
#define ECO_C3_NLEAF 21
#define ECO_C3_NACC 100 /* accumulated sums per shard (see csrc/eco_composite.cu for the layout) */

int64_t eco_composite3_ws_bytes(void);

/* acc_out: float64[ECO_C3_NACC], additive across shards (acc_out[0] = this shard's pixel count). */
int eco_composite3_stats(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits, void* ws,
                         int64_t ws_bytes, double* acc_out, int device, void* stream);

/* leaf scales: scale of each leaf's 7 losses, order = 3 channel leaves then, per pair
 * (0,1),(0,2),(1,2): I1,U1,I2,U2,I3,U3 (the host computes 2, 2, 2 and per pair 2*w_j, 2*w_i, 2*w_d, 2*w_i,
 * 2*w_d, 2*w_i*w_i*w_j with the numpy RNG stream of :49-52).  Given either as a host array (copied into
 * the launch) or, if leaf_scale_host is NULL, as a device array.
 * losses_out: float32[7]; jac_out: float64[21][7][ECO_NJAC]; leaf_sums_out: float64[21][ECO_NSTAT] (may be NULL). */
int eco_composite3_finalize(const double* acc, const double* leaf_scale_host, const double* leaf_scale_dev,
                            float* losses_out, double* jac_out, double* leaf_sums_out, int device, void* stream);

int eco_composite3_grad(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                        const double* jac, const float* upstream, const EcoOut* gx, int device, void* stream);

/* Single cooperative launch: stats -> grid barrier -> finalize -> gradient, for one GPU (no shard
 * exchange).  upstream is known up front (the weights of ess/train_multiclass.py:145). */
int eco_composite3_fused(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                         const double* leaf_scale_dev, const float* upstream, void* ws,
                         int64_t ws_bytes, float* losses_out, const EcoOut* gx, int device, void* stream);

/* The plain 3-organ multi-class loss step in ONE cooperative launch: ess/train_multiclass.py:253-274
 * `losses_fn(outputs, labels, composite_set_theory=False, ...)` for C == 3 (= the sum over channels of the 7-loss leaf
 * (a = g_c, b = x_c), :260-262; ess/loss_composite.py:28-40 is the same with leaf_scale = 2), with `F.sigmoid` (:134)
 * and `loss.backward()` (:147) fused: x are fp32 LOGITS with 16-byte aligned planes (H*W % 4 == 0; -8 otherwise, use
 * eco_pair_*), gx = d(sum_k upstream[k] * loss_k)/d logits, losses_out = float32[7] totals over the channels.
 * ws: eco_composite3_ws_bytes() bytes, zeroed once, not shared with a concurrent eco_composite3_* call. */
int eco_multiclass3_fused(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, double leaf_scale,
                          const float* upstream, void* ws, int64_t ws_bytes, float* losses_out, const EcoOut* gx,
                          int device, void* stream);
/* General form of the call above (which is this one with flags = 0 and upstream_prev = NULL):
 *   - flags & ECO_C3_PROBS: x holds probabilities instead of logits and gx = d(...)/d probabilities -- the reference's own
 *     call order, `outputs = F.sigmoid(net(x))` at ess/train_multiclass.py:134 BEFORE `losses_fn(outputs, labels,
 *     composite_set_theory=False, ...)` at :139-141, so an unchanged training loop runs on this kernel;
 *   - upstream_prev (device float32[7] or NULL): "only if changed" -- losses_out / gx already hold the step for the weights
 *     upstream_prev; the kernel compares the two vectors on the device and returns at once when they are equal (the
 *     backward half of the drop-in autograd path: the forward anticipated the weights of ess/train_multiclass.py:145). */
int eco_multiclass3_step(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, double leaf_scale,
                         const float* upstream, const float* upstream_prev, uint32_t flags, void* ws, int64_t ws_bytes,
                         float* losses_out, const EcoOut* gx, int device, void* stream);

/* Sharded flavour of eco_composite3_fused: one process per GPU, each with its batch shard; the 100 sums are
 * all-reduced INSIDE the kernel over NVLink peer memory (P2P stores + release/acquire flags), so the whole
 * data-parallel step -- ess/train_multiclass.py:133-147 on a sharded batch -- stays one launch per rank.
 * peer_xch_dev: device array of `world` pointers to every rank's exchange buffer as mapped into THIS process
 * (entry [rank] = own buffer).  epoch: 1, 2, 3, ... incremented by the caller on every call, identical on all
 * ranks.  Every rank must make the same sequence of calls (as with any collective). */
int eco_composite3_fused_sharded(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, int32_t from_logits,
                                 const double* leaf_scale_dev, const float* upstream, void* ws, int64_t ws_bytes,
                                 float* losses_out, const EcoOut* gx, void* const* peer_xch_dev, int32_t rank,
                                 int32_t world, uint32_t epoch, int device, void* stream);

/* The composite step, general form (the two calls above are this one with flags = from_logits ? 0 : ECO_C3_PROBS and no /
 * one peer exchange): ess/train_multiclass.py:133-147 for 3 organs with the composite loss of ess/loss_composite.py:21-94.
 *   - g may be ECO_U8: the datasets produce {0,1} masks (ess/dataset/fish/fish_dataset.py:159-171) that the reference
 *     converts to float before the loss (ess/train_multiclass.py:119-123); as bytes they cost 1 B instead of 4 B per
 *     element on the bus and in HBM.  Needs fp32 logits, 16-byte aligned planes and H*W % 16 == 0 (-8 otherwise);
 *   - ECO_C3_UNION_LABELS: g holds the RAW per-organ masks and the label union of ess/utils/subsets_union.py:8-32
 *     `return_union_sets_descending_order(ann, exclude_indices=[0])` (channel 1 <- channel 1 + channel 2, then everything
 *     above 1 set to 1; called on the labels at ess/train_multiclass.py:110) is applied in registers at load -- g itself is
 *     not modified, the separate in-place sweep (eco_union_sets) is not needed;
 *   - ECO_C3_PROBS: x holds probabilities (gradient w.r.t. them) instead of logits -- the reference's own call order,
 *     F.sigmoid at ess/train_multiclass.py:134 and then losses_fn on its output; fp32 with 16-byte aligned planes runs the
 *     same one-launch kernel as logits do;
 *   - ECO_C3_NO_GRAD: loss values only (losses_fn under torch.no_grad(), as in the validation loop of ess/train_multiclass.py:175-198): the
 *     statistics pass and the closed forms of the same kernel, no gradient pass; gx may be NULL.  Needs inputs the
 *     one-launch kernel serves (-8 otherwise: use eco_composite3_stats + eco_composite3_finalize);
 *   - peers == NULL: one GPU.  Otherwise the batch is sharded over peers->world processes and the per-class partial sums
 *     are all-reduced inside the kernel over NVLink peer memory (see eco_composite3_fused_sharded for the contract).
 *     A wait on a peer that exceeds timeout_ms (<= 0: 30 s) poisons the step's outputs with NaN and sets *status to 1:
 *     `status` is any device-visible word -- e.g. mapped pinned host memory, which the host can read without a
 *     synchronisation -- or NULL for a word inside `ws` (read and cleared by eco_xch_poll_status). */
#define ECO_C3_UNION_LABELS 1u
#define ECO_C3_PROBS 2u
#define ECO_C3_NO_GRAD 4u
typedef struct EcoPeerExchange {
    void* const* peer_xch_dev; /* device array of `world` exchange-buffer pointers, [rank] = own */
    int32_t rank;
    int32_t world;
    uint32_t epoch;            /* 1, 2, 3, ... identical on all ranks */
    uint32_t _pad;
    uint32_t* status;          /* device-visible word or NULL */
    double timeout_ms;         /* <= 0: default */
} EcoPeerExchange;
int eco_composite3_step(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, uint32_t flags,
                        const double* leaf_scale_dev, const float* upstream, void* ws, int64_t ws_bytes,
                        float* losses_out, const EcoOut* gx, const EcoPeerExchange* peers, int device, void* stream);
/* The same step, run only if the weights changed: `losses_out` / `gx` already hold the step for `upstream_prev` (an
 * earlier eco_composite3_step on the same inputs); the kernel compares the two float32[7] vectors ON THE DEVICE and returns at
 * once when they are equal, otherwise it recomputes the step for `upstream`.  This is how the drop-in `losses_fn(...)` ->
 * `loss.backward()` of ess/train_multiclass.py:139-147 becomes one launch in the forward (with the weights of :145 anticipated
 * from the previous step -- they depend on the epoch only, :92-100) and one empty launch in the backward, without a host
 * synchronisation to look at the weights.  Single GPU; fp32 logits with 16-byte aligned planes (-8 otherwise). */
int eco_composite3_step_if_changed(const EcoView* x, const EcoView* g, int32_t N, int64_t HW, uint32_t flags,
                                   const double* leaf_scale_dev, const float* upstream, const float* upstream_prev,
                                   void* ws, int64_t ws_bytes, float* losses_out, const EcoOut* gx, int device,
                                   void* stream);
/* Reads (and clears, when set) the time-out word inside `ws`; synchronises `stream`. */
int eco_xch_poll_status(void* ws, int64_t ws_bytes, uint32_t* status_out_host, int device, void* stream);

/* Exchange buffers for the call above (CUDA IPC needs a cudaMalloc base pointer, so this is the one place the
 * library allocates).  alloc: zeroed buffer of eco_xch_bytes(world) + its 64-byte IPC handle to send to the peers;
 * open/close: map / unmap a peer's buffer from its handle; free: release the own buffer. */
int64_t eco_xch_bytes(int32_t world);
int eco_xch_alloc(int32_t world, void** ptr_out, unsigned char* handle_out, int device);
int eco_xch_open(const unsigned char* handle, void** ptr_out, int device);
int eco_xch_close(void* peer_ptr, int device);
int eco_xch_free(void* ptr, int device);

/* ------------------------------------------------------------------------------------------
 * Evaluation scoring: ess/test_multiclass.py:58 (sigmoid), :68-69 (threshold rule, strict '>' in fp32),
 * :80-81 (per-class dice_loss(out_c, lab_c, background_weight=0)).
 * counts_out: int64[C][3] = (sum out*lab, sum out, sum lab) over pixels with out = (sigmoid(z) > T) and
 *             lab counted where it is exactly 1 (exact integers; see the _ex forms below for other label values);
 * soft_out:   float64[C][3] = (sum p*lab, sum p, sum lab^2) with p = sigmoid(z), the un-thresholded live path.
 * Both additive across shards.  thresholds: float32[n_thr] device (n_thr may be 0); counts_out is then
 * int64[n_thr][C][3] -- one read of logits+labels serves every threshold of the beam search (:64-77).
 * `logits_are_probs` is a flag word: ECO_EVAL_PROBS = the inputs already are probabilities (no sigmoid);
 * ECO_EVAL_UNUNION (soft Dice only, n_thr = 0; refused with -4 otherwise) = the prediction un-union of the sequential
 * model's test, ess/test_multiclass_sequential_densenetloss.py:66 -> ess/utils/subsets_union.py:21-27
 * (`return_union_sets_descending_order(out, reverse=True)`, exclude_indices=[0]: p_c <- |p_c - p_{c+1}| for
 * c = C-2 .. 1, in that order), taken in registers at load instead of a separate in-place sweep (eco_union_sets);
 * the predictions themselves are left untouched.
 * ------------------------------------------------------------------------------------------ */
#define ECO_EVAL_PROBS 1
#define ECO_EVAL_UNUNION 2
int64_t eco_dice_ws_bytes(int32_t C, int32_t n_thr);
int eco_dice_counts(const EcoView* logits, const EcoView* labels, int32_t N, int32_t C, int64_t HW,
                    const float* thresholds, int32_t n_thr, int32_t logits_are_probs, void* ws, int64_t ws_bytes,
                    int64_t* counts_out, double* soft_out, int device, void* stream);
/* Labels other than exactly 0 / 1 (masks resized by the dataset, ess/dataset/fish/fish_suim.py:60-74): the reference's
 * thresholded Dice is 2 sum(out*lab) / (sum out + sum lab^2) with the real label values, which integer counts cannot
 * express.  The `_ex` forms carry the per-threshold intersections as float64 next to the counts:
 *   thr_inter_out: float64[n_thr][C] = sum out*lab -- equal to the integer count for 0/1 labels; for a class with any other
 *   label value a second, exact pass over that class fills it (the counting kernel detects the case: a label is counted
 *   only where it is exactly 1, so sum lab^2 == count iff all labels are 0/1).  counts_out then holds (count of pixels
 *   with out = 1 and lab = 1, sum out, count of lab = 1) and eco_dice_finalize_ex evaluates the reference's formula from
 *   thr_inter, counts[.][1] and soft[.][2].  Without thr_inter_out (the plain forms) such a class gets counts_out[.][2] = -1
 *   and a NaN Dice from eco_dice_finalize -- loud, not silently different from the reference. */
int eco_dice_counts_ex(const EcoView* logits, const EcoView* labels, int32_t N, int32_t C, int64_t HW,
                       const float* thresholds, int32_t n_thr, int32_t logits_are_probs, void* ws, int64_t ws_bytes,
                       int64_t* counts_out, double* soft_out, double* thr_inter_out, int device, void* stream);
int eco_dice_finalize_ex(const int64_t* counts, const double* soft, const double* thr_inter, int32_t C, int32_t n_thr,
                         float* dice_out, float* soft_dice_out, int device, void* stream);
/* dice_out: float32[n_thr][C] thresholded and soft_dice_out: float32[C], each (2I+eps)/(U+eps); either may be NULL. */
int eco_dice_finalize(const int64_t* counts, const double* soft, int32_t C, int32_t n_thr, float* dice_out,
                      float* soft_dice_out, int device, void* stream);

/* Frame pre-processing in front of the network: ess/test_video.py:70-78
 *     transforms.Resize((256, 256)) -> transforms.ToTensor() -> transforms.Normalize(mean, std)
 * on RGB frames.  The reference resizes on the host with Pillow (Image.resize, BILINEAR: two 8-bit passes with 22-bit
 * fixed-point coefficients, horizontal first), divides by 255 and normalises in float32, then copies 4 B/element to the
 * GPU; here the uint8 frames go to the device as they are and one kernel produces the float32 [N][3][Hout][Wout] tensor,
 * bit-identical to Pillow + torchvision on the CPU.
 *   eco_frames_plan_sizes: taps per output column / row (ksx, ksy) for the table sizes below.
 *   eco_frames_plan (host only, no GPU): fills HOST arrays xbounds int32[Wout][2] = (first input column, tap count),
 *     kx int32[Wout][ksx], ybounds int32[Hout][2], ky int32[Hout][ksy] (Pillow's precompute_coeffs +
 *     normalize_coeffs_8bpc), lut float32[3][256] = ((byte / 255) - mean[c]) / std[c], and the largest input patch of
 *     one 32 x 16 output tile (patch_cols, patch_rows).  The caller copies the five arrays to the device once per size.
 *   eco_frames_preprocess: frames = uint8 device [N][Hin][Win][3] with the given byte strides between frames / rows;
 *     out = float32 device [N][3][Hout][Wout] contiguous.  -8 when one tile's input patch exceeds the shared memory
 *     (down-scaling by more than ~40x). */
int eco_frames_plan_sizes(int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, int32_t* ksx, int32_t* ksy);
int eco_frames_plan(int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, const float* mean, const float* stdev,
                    int32_t* xbounds, int32_t* kx, int32_t* ybounds, int32_t* ky, float* lut, int32_t* patch_cols,
                    int32_t* patch_rows);
int eco_frames_preprocess(const uint8_t* frames, int32_t N, int32_t Hin, int32_t Win, int64_t frame_stride_bytes,
                          int64_t row_stride_bytes, const int32_t* xbounds_dev, const int32_t* kx_dev, int32_t ksx,
                          const int32_t* ybounds_dev, const int32_t* ky_dev, int32_t ksy, int32_t Hout, int32_t Wout,
                          int32_t patch_cols, int32_t patch_rows, const float* lut_dev, float* out, int device, void* stream);

/* Byte masks for the result dumps that follow the scoring: ess/test_multiclass.py:58 (sigmoid), :68-69 (optional
 * threshold rule), :90-92 `(t.numpy() * 255).astype(np.uint8)` for images / labels / outputs, ess/test_video.py:129-130.
 * out[n][c][i] = uint8(trunc(fp32(q * 255))) with q = x (x_is_prob) or sigmoid(x), then, if use_threshold,
 * q = 1 where q > threshold (or q == 1), else 0.  x: f32 / bf16 view; out: contiguous uint8 [N][C][HW].
 * One pass, 5 B/element; bit-identical to the reference ops on the same device for q in [0, 1]. */
int eco_masks_u8(const EcoView* x, int32_t N, int32_t C, int64_t HW, float threshold, int32_t use_threshold,
                 int32_t x_is_prob, uint8_t* out, int device, void* stream);

/* ------------------------------------------------------------------------------------------
 * Soft-label cross entropy over the channel dim: ess/loss_functions.py:29-30
 * `F.cross_entropy(pred, gt) + bw * F.cross_entropy(1-pred, 1-gt)` with float targets.
 * a = gt (targets), b = pred (softmax input).  sums_out: float64[2] = (sum_pix sum_c a_c*logsoftmax(b)_c,
 * same for (1-a, 1-b)); loss = -(s0 + bw*s1)/n_pix.  Gradients w.r.t. both slots.
 * ------------------------------------------------------------------------------------------ */
int64_t eco_softce_ws_bytes(void);
int eco_softce_stats(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, int32_t need_bg, void* ws,
                     int64_t ws_bytes, double* sums_out, int device, void* stream);
int eco_softce_grad(const EcoView* a, const EcoView* b, int32_t N, int32_t C, int64_t HW, double background_weight,
                    double n_pix_total, const float* upstream, const EcoOut* ga, const EcoOut* gb, int device,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Label union / un-union adjacent to the path, IN PLACE: ess/utils/subsets_union.py:8-32
 * `return_union_sets_descending_order(ann, exclude_indices=[0], reverse=False)` (along the class dim) and its
 * batch-dim twin ess/train_multiclass.py:32-45 that train() calls at :110.  The tensor is viewed as
 * [outer][K][inner] (inner contiguous, element strides stride_outer / stride_k); bit k of exclude_mask = index k is
 * left alone.  forward: entry k (not excluded, not last) <- sum of entries k..K-1, then everything > 1 is set to 1;
 * reverse: from the back, entry k <- |entry k - entry k+1|.  K <= 64.
 * ------------------------------------------------------------------------------------------ */
int eco_union_sets(void* data, int32_t dtype, int64_t outer, int32_t K, int64_t inner, int64_t stride_outer,
                   int64_t stride_k, uint64_t exclude_mask, int32_t reverse, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECOLOSS_H */
