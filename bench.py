#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: fused per-pixel loss forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

Workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[1] = "cfg2": 3-organ composite
multiclass loss (loss_composite.losses_fn, composite_set_theory=True) on synthetic logits/masks
54 x 3 x 256 x 256 fp32, loss = bce + generalized_dice + twersky + focal_dice, forward AND backward
(d loss / d logits).  One "step" = one pass of that path over one batch.

Prints ONE JSON line (rank 0).  `value` = Gpixel/s with inputs resident in HBM; `e2e` = the same metric
through the public API with pinned HOST buffers (H2D of logits+masks and D2H of the 7 losses inside the
timed region); `roofline` = algorithmic bytes (12 B/element) / kernel time against the measured HBM peak;
`cpu_baseline` = the CPU oracle port of the reference path timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHTS = dict(bce=1.0, generalized_dice=1.0, twersky=1.0, focal_dice=1.0)  # BASELINE.md cfg2 combination
N_BUFFER_SETS = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2 (default, the headline) / cfg4: fused composite loss fwd+bwd; cfg5: frame-stream Dice "
                         "scoring (sigmoid -> threshold 0.8 -> exact counts -> per-batch Dice), 64x3x512x512 per GPU")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: all-reduce of the sums inside the fused kernel over peer memory, or NCCL between kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None

    def _start_nvml(self):
        """The timed region of a default run is ~15 ms, shorter than one nvidia-smi period: read the same counters
        (SM clock, max SM clock, clocks-event reasons) through NVML from a thread every 2 ms instead."""
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        state = {"sm": [], "reasons": set(), "stop": False, "max": mx}

        def pump():
            while not state["stop"]:
                try:
                    state["sm"].append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for nm, bit in bits.items():
                        if r & bit:
                            state["reasons"].add(nm)
                except Exception:
                    pass
                time.sleep(0.002)

        state["thread"] = threading.Thread(target=pump, daemon=True)
        state["thread"].start()
        self.nvml = state

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.nvml["stop"] = True
            self.nvml["thread"].join(timeout=1)
            sm = sorted(self.nvml["sm"])
            if sm:
                return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.nvml["max"], "reasons": sorted(self.nvml["reasons"]),
                        "samples": len(sm), "source": "nvml polled from a thread, warm-up + timed region"}
            return None
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------
def workload_shape(name):
    from ecologysemanticsegmentation_b200.synthetic import CONFIGS
    seed, n, c, s = CONFIGS[name]
    if name == "cfg4":
        n = 54  # per-GPU shard of the 432-image batch
    return seed, n, c, s


STREAM_THRESHOLD = 0.8   # cfg5: the first threshold of the reference's beam (test_multiclass.py:64)


def reference_impl():
    """("reference", lf, lc) = the UNMODIFIED reference modules staged under oracle/_ref/ by oracle/make_ref.py (they
    travel to the GPU box; /root/reference does not), else ("port", None, None) = the op-for-op restatement
    oracle/torch_port.py (bit-identical to the reference on CPU, tests/test_oracle_vs_reference.py)."""
    try:
        from oracle import make_ref
        if make_ref.staged():
            lf, lc = make_ref.load()
            return "reference", lf, lc
    except Exception as exc:   # fall back to the port, loudly
        print("oracle/_ref not usable, timing the port:", repr(exc), file=sys.stderr)
    return "port", None, None


def reference_kind_text(kind):
    return ("the reference's own loss_composite.losses_fn / loss_functions.dice_loss, unmodified (staged under oracle/_ref/)"
            if kind == "reference" else "oracle/torch_port.py (op-for-op restatement of the reference's eager path)")


def cpu_reference_step(z, g, weights, scoring=False):
    """The reference path on the tensors' device (the CPU for the baseline legs): sigmoid (train_multiclass.py:134) ->
    losses_fn(composite) -> weighted sum (:145) -> backward (:147); ``scoring``: one batch of test()'s scoring
    (sigmoid -> threshold rule -> per-class dice_loss, test_multiclass.py:58,68-69,80-82)."""
    import numpy as np
    import torch
    from oracle import torch_port as tp
    kind, lf, lc = reference_impl()
    if scoring:
        with torch.no_grad():
            if kind == "reference":
                out = tp.threshold_inplace(torch.sigmoid(z), STREAM_THRESHOLD)   # test_multiclass.py:58,68-69 (two statements)
                return [float(-lf.dice_loss(out[:, c:c + 1], g[:, c:c + 1], background_weight=0)) for c in range(g.shape[1])], None
            return [float(v) for v in tp.eval_batch_dice(z, g, STREAM_THRESHOLD)], None
    zz = z.clone().requires_grad_(True)
    np.random.seed(0)
    comp = z.shape[1] == 3
    if kind == "reference":
        losses = lc.losses_fn(torch.sigmoid(zz), g, comp)
    else:
        losses = tp.losses_composite(torch.sigmoid(zz), g, comp)
    total = sum(w * l for w, l in zip(weights, losses) if w != 0.0)
    total.backward()
    return [float(l) for l in losses], zz.grad


def config_dict(workload, world, exchange="p2p"):
    """The `config` object of the JSON line: ONE definition for both arms (ours / --impl reference), a function of
    the workload and the GPU count only."""
    seed, n, c, s = workload_shape(workload)
    if workload == "cfg5":
        return {"workload": f"cfg5: frame-stream Dice scoring (sigmoid -> threshold {STREAM_THRESHOLD} -> exact counts -> "
                            f"per-batch per-class Dice, mean over batches), {n}x{c}x{s}x{s} f32 logits + f32 masks per GPU and step",
                "global_batch": n * world,
                "parallelism": f"dp{world}: batches sharded over the ranks, counts all-reduced once per stream" if world > 1 else "single GPU",
                "l2": "GPU arm: rotating 2 batches of 403 MB (> 126 MB L2 each)"}
    return {"workload": f"{workload}: ORGANS=whole_body,ventral_side,dorsal_side composite multiclass loss "
                        f"fwd+bwd from logits, {n}x{c}x{s}x{s} f32 per GPU, loss=bce+gdice+twersky+focal_dice",
            "global_batch": n * world,
            "parallelism": f"dp{world}: batch sharded, per-class partial sums all-reduced" if world > 1 else "single GPU",
            "l2": f"GPU arm: rotating {N_BUFFER_SETS} buffer sets ({N_BUFFER_SETS * 3 * n * c * s * s * 4 / 1e6:.0f} MB) > 126 MB L2 between timed iterations"}


def time_cpu_baseline(name, budget_s=20.0, n_images=None, steps=None, warmup=1):
    """Bounded CPU sample: `n_images` images of the workload per step."""
    import torch
    from ecologysemanticsegmentation_b200 import fused
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    seed, n, c, s = workload_shape(name)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    weights = fused.loss_weights(**WEIGHTS)
    if n_images is None:
        n_images = n
    z, g = make_inputs(n_images, c, s, seed)
    scoring = name == "cfg5"
    for _ in range(warmup):
        cpu_reference_step(z, g, weights, scoring)
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        cpu_reference_step(z, g, weights, scoring)
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif len(times) >= 3 and (time.perf_counter() - t_start) > budget_s or len(times) >= 10:
            break
    pixels = n_images * s * s
    return {"times": times, "pixels": pixels, "cores": torch.get_num_threads(), "n_images": n_images, "shape": (n_images, c, s, s)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (its eager ops, restated op for op in
    oracle/torch_port.py because /root/reference does not travel to the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seed, n, c, s = workload_shape(args.workload)
    # size the per-step sample so the whole run ends within ~2 minutes
    probe = time_cpu_baseline(args.workload, n_images=2, steps=1, warmup=1)
    per_image = probe["times"][0] / 2
    total_steps = args.steps + args.warmup
    n_images = int(max(1, min(n, 100.0 / max(per_image * total_steps, 1e-9))))
    res = time_cpu_baseline(args.workload, n_images=n_images, steps=args.steps, warmup=args.warmup)
    t = sum(res["times"]) / len(res["times"])
    value = res["pixels"] / t / 1e9
    kind = reference_impl()[0]
    sample = (f"{n_images} of {n} images of {args.workload} ({n_images}x{c}x{s}x{s}) per step, {args.steps} steps, "
              f"{res['cores']} host threads; {reference_kind_text(kind)}")
    scoring = args.workload == "cfg5"
    line = {
        "impl": "reference", "metric": "Gpixel/s Dice eval" if scoring else "Gpixel/s fused loss fwd+bwd", "value": value, "unit": "Gpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": res["cores"], "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure_aux(dev):
    """Other kernels of the path on BASELINE.json's other single-GPU configs, for context next to the headline:
    cfg3 = per-class thresholded Dice scoring at 54x3x1024x1024 (8 B/element), cfg1 = single-class leaf fwd+bwd.
    Inputs are larger than L2 (cfg3: 1.36 GB) or rotated (cfg1)."""
    import torch
    from ecologysemanticsegmentation_b200 import ops
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(iters):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / iters * 1e-3

    def timed_graph(fn, per_replay, replays):
        """Short kernels (tens of us) are launch-path bound from Python: capture `per_replay` calls of `fn` (one per
        rotating buffer set) in a CUDA graph and time its replays.  Returns (seconds per call, "graph" | "eager")."""
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(per_replay):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(per_replay):
                    fn()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(replays):
                graph.replay()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / (replays * per_replay) * 1e-3, "graph"
        except Exception as exc:  # capture not possible on this driver: plain launches
            print("graph capture failed, timing plain launches:", repr(exc), file=sys.stderr)
            torch.cuda.synchronize()
            return timed(fn, replays * per_replay), "eager"

    n, c, s = 54, 3, 1024
    z, g = make_inputs(n, c, s, 103)
    z, g = z.to(dev), g.to(dev)
    for label, thr in (("soft", None), ("threshold_0.8", torch.tensor([0.8], dtype=torch.float32, device=dev))):
        t = timed(lambda: ops.dice_counts(z, g, thr), 20)
        gbs = 8.0 * n * c * s * s / t / 1e9
        out["dice_eval_cfg3_" + label] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs,
                                          "frac_of_hbm_peak": gbs / peak, "bytes_per_element": 8}
    # the sequential model's test (ess/test_multiclass_sequential_densenetloss.py:62,66,97-99): sigmoid -> prediction un-union
    # -> soft Dice, fused into the one scoring read, against the same result from the stand-alone steps
    from ecologysemanticsegmentation_b200 import subsets_union
    t = timed(lambda: ops.dice_counts_ex(z, g, None, ununion_preds=True), 20)
    pbuf = torch.empty_like(z)

    def unfused():
        torch.sigmoid(z, out=pbuf)
        subsets_union.return_union_sets_descending_order(pbuf, reverse=True)
        ops.dice_counts_ex(pbuf, g, None, inputs_are_probs=True)
    t_unfused = timed(unfused, 10)
    del pbuf
    gbs = 8.0 * n * c * s * s / t / 1e9
    out["dice_eval_cfg3_soft_ununion"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs,
                                          "frac_of_hbm_peak": gbs / peak, "bytes_per_element": 8,
                                          "us_unfused": t_unfused * 1e6,
                                          "what": "sigmoid -> |p1 - p2| un-union -> soft Dice in ONE read (the channel-1 CTAs read "
                                                  "plane 2 as well: L2 hits); unfused = torch.sigmoid + the in-place un-union kernel "
                                                  "+ scoring of the probabilities"}
    # the threshold beam search of ess/test_multiclass.py:64-77 (np.arange(0.8, 0.99, 0.01): 19 thresholds) from ONE read
    import numpy as np
    thr19 = torch.tensor(np.arange(0.8, 0.99, step=0.01), dtype=torch.float32, device=dev)
    t = timed(lambda: ops.dice_counts(z, g, thr19), 20)
    gbs = 8.0 * n * c * s * s / t / 1e9
    out["dice_eval_cfg3_beam_19_thresholds"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs,
                                                "frac_of_hbm_peak": gbs / peak, "bytes_per_element": 8,
                                                "what": "all 19 thresholds of the beam search in one launch (binning kernel)"}
    del z, g
    # the plain 3-organ multi-class loss step (the loss train_multiclass.py trains with), cfg2's shape, one launch
    from ecologysemanticsegmentation_b200 import fused
    n, c, s = 54, 3, 256
    sets = [tuple(t.to(dev) for t in make_inputs(n, c, s, 102 + 7 * k)) for k in range(4)]
    outs = [torch.empty_like(zz) for zz, _ in sets]
    mstep = fused.MulticlassLossStep(fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0, focal_dice=1.0), device=dev)
    state = {"i": 0}

    def run_mc():
        k = state["i"] % 4
        state["i"] += 1
        mstep(sets[k][0], sets[k][1], out=outs[k])

    t, how = timed_graph(run_mc, 4, 25)
    gbs = 12.0 * n * c * s * s / t / 1e9
    out["multiclass_plain_cfg2_shape"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs,
                                          "frac_of_hbm_peak": gbs / peak, "bytes_per_element": 12, "launch": how,
                                          "what": "train_multiclass.losses_fn (3 plain leaves) fwd+bwd from logits, one launch"}
    # cfg1 (BASELINE.json configs[0]): ORGANS=whole_body single-class loss fwd+bwd from logits, one launch (eco_pair_fused)
    n, c, s = 54, 1, 256
    sets1 = [tuple(t.to(dev) for t in make_inputs(n, c, s, 101 + 7 * k)) for k in range(12)]   # 12 x 28 MB > L2
    outs1 = [torch.empty_like(zz) for zz, _ in sets1]
    lstep = fused.LeafLossStep(fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0), doubling=1.0, device=dev)
    state1 = {"i": 0}

    def run_leaf():
        k = state1["i"] % 12
        state1["i"] += 1
        lstep(sets1[k][0], sets1[k][1], out=outs1[k])

    t, how = timed_graph(run_leaf, 12, 20)
    gbs = 12.0 * n * c * s * s / t / 1e9
    out["leaf_cfg1"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak,
                        "bytes_per_element": 12, "launch": how,
                        "what": "cfg1: single-class losses_fn (bce + gdice + twersky) fwd+bwd from logits, 54x1x256x256, one launch"}
    del sets1, outs1
    # ---- the composite step on the other shapes / input forms ------------------------------------------------------
    import ecologysemanticsegmentation_b200 as eco
    w = fused.loss_weights(**WEIGHTS)
    np.random.seed(0)
    cstep = fused.CompositeLossStep(w, device=dev)
    # (a) byte masks on the device: 9 instead of 12 B/element
    n, c, s = 54, 3, 256
    sets8 = [(zz, gg.to(torch.uint8)) for zz, gg in sets]
    st8 = {"i": 0}

    def run_u8():
        k = st8["i"] % 4
        st8["i"] += 1
        cstep(sets8[k][0], sets8[k][1], out=outs[k])

    t, how = timed_graph(run_u8, 4, 25)
    out["composite_cfg2_u8_masks"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": 9.0 * n * c * s * s / t / 1e9,
                                      "frac_of_hbm_peak": 9.0 * n * c * s * s / t / 1e9 / peak, "bytes_per_element": 9, "launch": how,
                                      "what": "cfg2 with uint8 masks (ECO_U8), one launch"}
    # (a2) probabilities in (ECO_C3_PROBS: the reference's own call order, F.sigmoid before losses_fn) and the step without
    #      gradient (ECO_C3_NO_GRAD: losses_fn under torch.no_grad()), same kernel
    psets = [torch.sigmoid(zz) for zz, _ in sets]
    pstep = fused.CompositeLossStep(w, device=dev, from_logits=False)

    def run_probs():
        k = st8["i"] % 4
        st8["i"] += 1
        pstep(psets[k], sets[k][1], out=outs[k])

    t, how = timed_graph(run_probs, 4, 25)
    out["composite_cfg2_probabilities"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": 12.0 * n * c * s * s / t / 1e9,
                                           "frac_of_hbm_peak": 12.0 * n * c * s * s / t / 1e9 / peak, "bytes_per_element": 12, "launch": how,
                                           "what": "cfg2 from probabilities (gradient w.r.t. them), one launch"}
    ents = [ops.PreparedComposite3(sets[k][0], sets[k][1], cstep.scales, cstep.upstream, True) for k in range(4)]

    def run_nograd():
        k = st8["i"] % 4
        st8["i"] += 1
        ents[k].run(no_grad=True)

    t, how = timed_graph(run_nograd, 4, 25)
    out["composite_cfg2_no_grad"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": 8.0 * n * c * s * s / t / 1e9,
                                     "frac_of_hbm_peak": 8.0 * n * c * s * s / t / 1e9 / peak, "bytes_per_element": 8, "launch": how,
                                     "what": "cfg2 loss values only (validation, no gradient pass), one launch"}
    del ents, psets
    # (b) the drop-in path through the reference's own signature: losses_fn(...) -> weighted sum -> backward(), exactly as
    #     ess/train_multiclass.py:139-147 calls it (from_logits=True fuses the sigmoid of :134); eager launches
    zs = [zz.clone().requires_grad_(True) for zz, _ in sets]

    def run_dropin():
        k = st8["i"] % 4
        st8["i"] += 1
        zs[k].grad = None
        ce, bce, fl, dice, gdice, tw, fd = eco.losses_fn(zs[k], sets[k][1], True, from_logits=True)
        loss = 1.0 * fd + 1.0 * bce + 1.0 * (gdice + tw)      # :145 with all three weights at 1
        loss.backward()

    t = timed(run_dropin, 40)
    out["dropin_autograd_cfg2"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "launch": "eager (autograd): host-bound, "
                                   "~25 python-level torch calls per step",
                                   "what": "eco.losses_fn(z, g, True, from_logits=True) -> weighted sum -> .backward(), cfg2"}
    # the same step (forward through the reference signature + autograd backward) captured in a CUDA graph, the way a
    # training loop that is launch-bound from python would run it: what the GPU needs for the drop-in path
    try:
        zg = sets[0][0].clone().requires_grad_(True)
        gg = sets[0][1]

        def dropin_step():
            ce, bce, fl, dice, gdice, tw, fd = eco.losses_fn(zg, gg, True, from_logits=True)
            loss = 1.0 * fd + 1.0 * bce + 1.0 * (gdice + tw)
            loss.backward()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                zg.grad = None
                dropin_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        zg.grad = None
        with torch.cuda.graph(graph, stream=side):
            dropin_step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        ref_grad = zg.grad.clone()
        ev0.record()
        for _ in range(100):
            graph.replay()
        ev1.record()
        torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1) / 100 * 1e-3
        zz = sets[0][0].clone().requires_grad_(True)
        l2 = eco.losses_fn(zz, gg, True, from_logits=True)
        (1.0 * l2[6] + 1.0 * l2[1] + 1.0 * (l2[4] + l2[5])).backward()
        same = bool(torch.equal(zz.grad, ref_grad))
        out["dropin_autograd_cfg2_graphed"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "launch": "one CUDA graph per step (forward + autograd backward captured)",
                                               "grad_equals_eager": same,
                                               "what": "the same drop-in step replayed as a CUDA graph"}
        del graph
    except Exception as exc:
        out["dropin_autograd_cfg2_graphed"] = {"error": repr(exc)[:300]}
    # (b2) the reference's training step with NOTHING changed but the import: outputs = F.sigmoid(outputs) (:134), then
    #      losses_fn on the probabilities (:139), weighted sum (:145), backward (:147); captured in a CUDA graph
    try:
        zg = sets[0][0].clone().requires_grad_(True)
        gg = sets[0][1]

        def unchanged_step():
            outputs = torch.sigmoid(zg)
            ce, bce, fl, dice, gdice, tw, fd = eco.losses_fn(outputs, gg, True)
            loss = 1.0 * fd + 1.0 * bce + 1.0 * (gdice + tw)
            loss.backward()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                zg.grad = None
                unchanged_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        zg.grad = None
        with torch.cuda.graph(graph, stream=side):
            unchanged_step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(100):
            graph.replay()
        ev1.record()
        torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1) / 100 * 1e-3
        out["dropin_unchanged_train_loop_cfg2_graphed"] = {
            "gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "launch": "one CUDA graph per step",
            "what": "F.sigmoid(z) -> eco.losses_fn(probabilities, g, True) -> weighted sum -> backward: torch's sigmoid forward "
                    "and backward kernels around the one-launch composite step on probabilities"}
        del graph
    except Exception as exc:
        out["dropin_unchanged_train_loop_cfg2_graphed"] = {"error": repr(exc)[:300]}
    # (b3) the call the reference's train() REALLY makes (ess/train_multiclass.py:134,139-141,145,147: composite_set_theory is
    #      hard-wired to False there): F.sigmoid -> the plain losses_fn on the probabilities -> weighted sum -> backward, with
    #      nothing changed but the import -- for three organs (cfg2's shape) and for ORGANS=whole_body (cfg1, the reference's
    #      default): one launch of the fused plain / leaf step on probabilities with anticipated weights (+ the backward's
    #      "only if changed" check) against the three pair-leaf launches the same call took before
    from ecologysemanticsegmentation_b200 import train_multiclass as tmod

    def live_loop(zsrc, gsrc, with_fd):
        zg = zsrc.clone().requires_grad_(True)

        def live_step():
            outputs = torch.sigmoid(zg)
            ce, bce, fl, dice, gdice, tw, fd = tmod.losses_fn(outputs, gsrc, composite_set_theory=False, background_weight=0,
                                                              early_stopped=False)
            loss = 1.0 * bce + 1.0 * (gdice + tw)          # :145 with the epoch < 1000 weights (:92-100) ...
            if with_fd:
                loss = 1.0 * fd + loss                      # ... and with focal_dice_w = 1
            loss.backward()

        def graphed_us():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    zg.grad = None
                    live_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            zg.grad = None
            with torch.cuda.graph(graph, stream=side):
                live_step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(100):
                graph.replay()
            ev1.record()
            torch.cuda.synchronize()
            del graph
            return ev0.elapsed_time(ev1) / 100 * 1e3

        us_fast = graphed_us()
        grad_fast = zg.grad.clone()
        ops.PLAIN_FAST_PATH = False
        try:
            us_three = graphed_us()
        finally:
            ops.PLAIN_FAST_PATH = True
        err = float((grad_fast - zg.grad).abs().max() / zg.grad.abs().max())
        return us_fast, us_three, err

    for name, zsrc, gsrc, with_fd, what in (
            ("dropin_live_train_loop_plain_cfg2_shape_graphed", sets[0][0], sets[0][1], True,
             "54x3x256x256, three organs: ONE launch of the plain fused step on probabilities"),
            ("dropin_live_train_loop_cfg1_graphed", sets[0][0][:, :1].contiguous(), sets[0][1][:, :1].contiguous(), False,
             "cfg1 (ORGANS=whole_body, 54x1x256x256, loss = bce + gdice + twersky): ONE launch of the resident leaf step on "
             "probabilities")):
        try:
            us_fast, us_three, err = live_loop(zsrc, gsrc, with_fd)
            out[name] = {"gpixel_per_s": n * s * s / (us_fast * 1e-6) / 1e9, "us": us_fast, "us_three_launch_path": us_three,
                         "grad_vs_three_launch_path_maxnorm": err, "launch": "one CUDA graph per step",
                         "what": "F.sigmoid(z) -> train_multiclass.losses_fn(probabilities, g, composite_set_theory=False) -> "
                                 "weighted sum -> backward; " + what + "; torch's sigmoid forward / backward around it; before: "
                                 "statistics + closed forms + gradient launches of the pair-leaf kernels"}
        except Exception as exc:
            out[name] = {"error": repr(exc)[:300]}
    # (c) the reference's own eager ops on THIS GPU (the like-for-like 'before'): same step, cfg2, oracle port on cuda
    try:
        t = timed(lambda: cpu_reference_step(sets[0][0], sets[0][1], w), 3)
        out["reference_eager_same_gpu_cfg2"] = {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6,
                                                "what": "the reference path (sigmoid -> losses_fn composite -> backward) in eager torch ops on this GPU: "
                                                        + reference_kind_text(reference_impl()[0])}
    except Exception as exc:
        out["reference_eager_same_gpu_cfg2"] = {"error": repr(exc)}
    del sets8, zs, sets, outs
    torch.cuda.empty_cache()
    # (e) frame pre-processing in front of the network (ess/test_video.py:70-78; upstream of the path, BASELINE configs[4]'s
    #     "1080p frames resized to 512"): uint8 frames on the device -> normalised float32 batch, one kernel
    try:
        from ecologysemanticsegmentation_b200 import test_video as tv
        nf = 64
        gen = torch.Generator(device="cpu").manual_seed(105)
        frames = torch.randint(0, 256, (nf, 1080, 1920, 3), dtype=torch.uint8, generator=gen).to(dev)
        fout = torch.empty((nf, 3, 512, 512), dtype=torch.float32, device=dev)
        t = timed(lambda: tv.preprocess_frames(frames, (512, 512), out=fout), 10)
        nbytes = frames.numel() + 4.0 * fout.numel()
        out["frames_preprocess_1080p_to_512"] = {"frames_per_s": nf / t, "us": t * 1e6, "gb_per_s": nbytes / t / 1e9,
                                                 "frac_of_hbm_peak": nbytes / t / 1e9 / peak,
                                                 "bytes_per_frame": nbytes / nf,
                                                 "what": "64 uint8 1080p frames -> Pillow-exact bilinear resize to 512x512 + ToTensor + Normalize, "
                                                         "float32 [64,3,512,512], one launch (3 B per input pixel + 12 B per output pixel)"}
        del frames, fout
        torch.cuda.empty_cache()
    except Exception as exc:
        out["frames_preprocess_1080p_to_512"] = {"error": repr(exc)[:300]}
    # (d) cfg4's per-GPU shard (BASELINE configs[3]: 432x3x512x512 over 8 GPUs = 54x3x512x512 per GPU), single GPU
    out["composite_cfg4_shard"] = time_cfg4_shard(dev, cstep, peak)
    return out


def time_cfg4_shard(dev, step, peak, nsets=2, iters=40):
    """us per step of `step` on 54x3x512x512 fp32 (cfg4's per-GPU shard), rotating `nsets` buffer sets (340 MB each, > L2)."""
    import torch
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    n, c, s = 54, 3, 512
    sets = [tuple(t.to(dev) for t in make_inputs(n, c, s, 104 + 7 * k)) for k in range(nsets)]
    outs = [torch.empty_like(zz) for zz, _ in sets]
    for i in range(3):
        step(sets[i % nsets][0], sets[i % nsets][1], out=outs[i % nsets])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(iters):
        step(sets[i % nsets][0], sets[i % nsets][1], out=outs[i % nsets])
    ev1.record()
    torch.cuda.synchronize()
    t = ev0.elapsed_time(ev1) / iters * 1e-3
    gbs = 12.0 * n * c * s * s / t / 1e9
    del sets, outs
    torch.cuda.empty_cache()
    return {"gpixel_per_s": n * s * s / t / 1e9, "us": t * 1e6, "gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak,
            "bytes_per_element": 12, "what": "cfg4 per-GPU shard: composite loss fwd+bwd, 54x3x512x512 f32, one launch"}



# ----------------------------------------------------------------------------------------------------
# parity checks printed with the line (outside the timed region); a failure exits non-zero
# ----------------------------------------------------------------------------------------------------
def _gather_batch(t, world):
    """All ranks' tensors concatenated along the batch, in rank order, on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


def _max_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def composite_parity(step, z, g, weights, dev, world, rank):
    """the sharded step must equal the single-device step on the concatenated batch.  Every rank
    gathers the global batch, runs the single-GPU fused step on it and compares the 7 losses (relative) and its own
    gradient shard (max-norm, relative to the global max) with what the sharded step returned for the same inputs.
    At N=1 the same check runs against the ORACLE on this device (reference semantics, autograd) at the full size."""
    import numpy as np
    import torch
    from ecologysemanticsegmentation_b200 import fused
    out = {}
    losses, grad = step(z, g)
    torch.cuda.synchronize()
    if world > 1:
        zf, gf = _gather_batch(z, world), _gather_batch(g, world)
        np.random.seed(0)
        single = fused.CompositeLossStep(weights, device=dev)
        lf, gradf = single(zf, gf)
        torch.cuda.synchronize()
        n = z.shape[0]
        mine = gradf[rank * n:(rank + 1) * n]
        e_loss = float(((losses[1:] - lf[1:]).abs() / lf[1:].abs()).max())
        e_grad = float((grad - mine).abs().max() / gradf.abs().max())
        nan = float(not (bool(torch.isfinite(losses).all()) and bool(torch.isfinite(grad).all())))
        e_loss, e_grad, nan = _max_over_ranks([e_loss, e_grad, nan], dev, world)
        out = {"vs": f"single-GPU fused step on the gathered global batch ({zf.shape[0]}x{zf.shape[1]}x{zf.shape[2]}x{zf.shape[3]}), max over ranks",
               "loss_rel": e_loss, "grad_maxnorm": e_grad, "tol": 1e-6, "ok": bool(e_loss <= 1e-6 and e_grad <= 1e-6 and nan == 0.0)}
        del zf, gf, gradf
    else:
        from oracle import torch_port as tp
        zr = z.clone().requires_grad_(True)
        np.random.seed(0)
        ref = tp.losses_composite(torch.sigmoid(zr), g, True)
        sum(w * l for w, l in zip(weights, ref) if w != 0.0).backward()
        rl = torch.stack([v.detach() for v in ref]).double()
        e_loss = float(((losses.double()[1:] - rl[1:]).abs() / rl[1:].abs()).max())
        d = (grad - zr.grad).double()
        e_max = float(d.abs().max() / zr.grad.abs().max())
        e_l2 = float(d.norm() / zr.grad.double().norm())
        out = {"vs": "oracle/torch_port.py (reference semantics, eager autograd) on the same device, same inputs, full size",
               "loss_rel": e_loss, "grad_maxnorm": e_max, "grad_rel_l2": e_l2, "tol": 1e-5,
               "ok": bool(e_loss <= 1e-5 and e_max <= 1e-5 and e_l2 <= 1e-5 and float(losses[0]) == 0.0)}
        del zr, ref
    torch.cuda.empty_cache()
    return out


def stream_parity(z, g, dev, world, group):
    """cfg5: the sharded scorer's counts for one batch must equal (integer-exactly) the counts of the gathered global
    batch scored on one GPU; at N=1 the counts are compared with the oracle's exact-integer counter on this device."""
    import torch
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    d_sh, c_sh, _ = tmc.score_batch(z, g, STREAM_THRESHOLD, group=group, return_counts=True)
    if world > 1:
        zf, gf = _gather_batch(z, world), _gather_batch(g, world)
        d_f, c_f, _ = tmc.score_batch(zf, gf, STREAM_THRESHOLD, return_counts=True)
        bad = float(not (torch.equal(c_sh, c_f) and torch.equal(d_sh, d_f)))
        bad = _max_over_ranks([bad], dev, world)[0]
        return {"vs": f"single-GPU scoring of the gathered global batch ({zf.shape[0]} images)", "counts_equal": bad == 0.0, "ok": bad == 0.0}
    from oracle import counts as oc
    ref = oc.batch_counts(z, g, STREAM_THRESHOLD)
    ok = bool((c_sh[0].cpu().numpy() == ref).all())
    return {"vs": "oracle/counts.py (reference threshold rule + int64 sums) on the same device", "counts_equal": ok, "ok": ok}


def bind_host_to_gpu_numa_node(device_index):
    """Pinned host buffers are allocated on the NUMA node of the allocating thread (first touch): run this process on the
    cores next to its GPU so that every rank's H2D stream leaves from local memory.  Returns a note for the JSON line."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        if hasattr(props, "pci_bus_id"):
            bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        else:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device_index
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(phys)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]   # nvml prints an 8-digit domain, sysfs a 4-digit one
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa node of the GPU unknown"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"numa node {node}: none of its cores is available to this process"
        os.sched_setaffinity(0, allowed)
        return f"process bound to the {len(allowed)} cores of NUMA node {node} (GPU {bus})"
    except Exception as exc:
        return "no NUMA binding: " + repr(exc)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # Everything but the one JSON line goes to stderr: libraries (NCCL's version banner, for one) write to fd 1.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run(args)
    finally:
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)
        par = line.get("parity")
        if isinstance(par, dict) and par.get("ok") is False:
            print("PARITY FAILED:", json.dumps(par), file=sys.stderr)
            sys.exit(3)


def _run_stream(args):
    """--workload cfg5 (BASELINE.json configs[4]): the frame stream of test_video / the batch loop of test_multiclass.
    One step = one batch of 64x3x512x512 logits + masks per GPU scored by ONE launch of the scoring kernel (sigmoid ->
    strict '>' 0.8 -> exact int64 counts + soft sums) into its slot of the stream buffer; the global batch is sharded
    over the ranks, and because the per-batch Dice is only needed at the end (mean over batches, test_multiclass.py:104)
    the whole [steps, C, 3] buffer is all-reduced ONCE, inside the timed region, followed by the closed forms."""
    import torch
    import torch.distributed as dist
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from ecologysemanticsegmentation_b200.synthetic import make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seed, n, c, s = workload_shape("cfg5")
    nsets = 2   # 2 x 403 MB per GPU: every batch is 3x larger than L2, nothing is L2-resident between steps
    host_sets = [make_inputs(n, c, s, seed + 1000 * rank + 17 * k, pin=True) for k in range(nsets)]
    dev_sets = [(z.to(dev), g.to(dev)) for z, g in host_sets]
    group = "world" if world > 1 else None
    warm = max(args.warmup, 3)
    scorer = tmc.StreamScorer(c, max(args.steps, warm), STREAM_THRESHOLD, device=dev, group=group)
    pixels_per_step = n * s * s * world
    elems_per_gpu = n * c * s * s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None   # from the warm-up on, through the timed region
    if sampler:
        sampler.start()
    for i in range(warm):
        scorer.add(*dev_sets[i % nsets])
    scorer.result()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scorer.reset()
    barrier()
    ev0.record()
    for i in range(args.steps):
        scorer.add(*dev_sets[i % nsets])
    dice = scorer.result()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        e_steps = max(3, min(args.steps, 20))
        zd, gd = torch.empty_like(dev_sets[0][0]), torch.empty_like(dev_sets[0][1])

        def stream_e2e(k):
            scorer.reset()
            for i in range(k):
                zh, gh = host_sets[i % nsets]
                zd.copy_(zh, non_blocking=True)
                gd.copy_(gh, non_blocking=True)
                scorer.add(zd, gd)
            return scorer.result().cpu()   # device -> host read of the stream's result (synchronises)

        stream_e2e(2)
        barrier()
        ev0.record()
        stream_e2e(e_steps)
        ev1.record()
        barrier()
        te = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item()) / e_steps
        e2e = {"value": pixels_per_step / (e_ms * 1e-3) / 1e9, "unit": "Gpixel/s",
               "h2d_bytes_per_step": 2 * elems_per_gpu * 4 * world, "d2h_bytes_per_step": c * 4 * world / e_steps,
               "ms_per_step": e_ms, "steps": e_steps}

    parity = stream_parity(dev_sets[0][0], dev_sets[0][1], dev, world, group)

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg_bytes = 8.0 * elems_per_gpu   # read logits 4 + read masks 4 per element; outputs are O(C)
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None   # dram__bytes_read + write per launch of the scoring kernel from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("cfg5")
            except Exception:
                traffic = None
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            res = time_cpu_baseline("cfg5", budget_s=15.0, n_images=16)
            tb = min(res["times"])
            kind = reference_impl()[0]
            cpu_baseline = {"value": res["pixels"] / tb / 1e9, "unit": "Gpixel/s", "cores": res["cores"], "kind": kind,
                            "sample": f"{len(res['times'])} batches of 16 of the 64 images of a cfg5 batch (16x{c}x{s}x{s}), best of; "
                                      + reference_kind_text(kind),
                            "ms_per_step": tb * 1e3}
        line = {
            "metric": "Gpixel/s Dice eval", "value": value, "unit": "Gpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict("cfg5", world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "dice_counts_kernel<float,float,4,1>", "algorithmic_bytes_per_launch": alg_bytes},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": (args.steps + 1) * world,
            "dice": [float(v) for v in dice.cpu()],
            "parity": parity,
        }
    if world > 1:
        dist.destroy_process_group()
    return line


def _run(args):
    if args.workload == "cfg5":
        return _run_stream(args)

    import torch
    import torch.distributed as dist
    from ecologysemanticsegmentation_b200 import fused
    from ecologysemanticsegmentation_b200.synthetic import make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_host_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    seed, n, c, s = workload_shape(args.workload)
    weights = fused.loss_weights(**WEIGHTS)
    import numpy as np
    np.random.seed(0)
    if world > 1 and args.exchange == "p2p":
        step = fused.PeerShardedCompositeLossStep(weights, group="world", device=dev)
    elif world > 1:
        step = fused.ShardedCompositeLossStep(weights, group="world", device=dev)
    else:
        step = fused.CompositeLossStep(weights, device=dev)

    # each rank owns its own 54-image shard (distinct seeds): weak scaling, global batch = 54 * world
    host_sets = []
    for k in range(N_BUFFER_SETS):
        z, g = make_inputs(n, c, s, seed + 1000 * rank + 17 * k, pin=True)
        host_sets.append((z, g))
    dev_sets = [(z.to(dev), g.to(dev)) for z, g in host_sets]
    grads = [torch.empty_like(z) for z, _ in dev_sets]
    pixels_per_step = n * s * s * world
    elems_per_gpu = n * c * s * s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(i):
        z, g = dev_sets[i % N_BUFFER_SETS]
        return step(z, g, out=grads[i % N_BUFFER_SETS])

    # clocks are sampled from the warm-up on (same kernel, same load) through the timed region: a default timed
    # region lasts ~15 ms, about one sampling period
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        one(i)
    barrier()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        losses, _ = one(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e9
    launches_per_step = 1 if (world == 1 or args.exchange == "p2p") else 3

    # ---- end to end: pinned host buffers in, 7 losses out, every step --------------------------------
    # The call a user of the step makes, fed from HOST memory: every step copies its logits (fp32) and its masks from
    # pinned buffers, runs the fused step and reads the 7 loss values back to the host.  The masks travel as uint8 -- the
    # datasets produce {0,1} masks (ess/dataset/fish/fish_dataset.py:159-171), the byte form is lossless and the kernel
    # takes it directly (ECO_U8): 53 instead of 85 MB per step.  Copies run on their own stream into two rotating device
    # buffer sets, so step i+1's input crosses PCIe while step i computes; the losses of step i are read one step late.
    # `e2e_f32_masks` is the same loop with fp32 masks and no overlap (copy -> copy -> step -> read), as in round 1.
    e2e = None
    e2e_variants = {}
    if not args.no_e2e:
        e_steps = max(3, min(args.steps, 50))
        host8 = [(z, g.to(torch.uint8).pin_memory()) for z, g in host_sets]

        def e2e_state(masks_u8, overlap):
            main = torch.cuda.current_stream()
            return {"zd": [torch.empty_like(dev_sets[0][0]) for _ in range(2)],
                    "gd": [torch.empty((n, c, s, s), dtype=torch.uint8 if masks_u8 else torch.float32, device=dev) for _ in range(2)],
                    "loss_host": [torch.empty(7, dtype=torch.float32).pin_memory() for _ in range(2)],
                    "main": main, "copy": torch.cuda.Stream() if overlap else main,
                    "ev": [[torch.cuda.Event() for _ in range(2)] for _ in range(3)]}

        def run_e2e(k, masks_u8, overlap, st):
            src = host8 if masks_u8 else host_sets
            zd, gd, loss_host, main, copy_stream = st["zd"], st["gd"], st["loss_host"], st["main"], st["copy"]
            ready, free, done = st["ev"]
            got = []
            for i in range(k):
                b = i % 2
                zh, gh = src[i % N_BUFFER_SETS]
                with torch.cuda.stream(copy_stream):
                    if overlap and i >= 2:
                        copy_stream.wait_event(free[b])      # the step that last used this buffer set has finished
                    zd[b].copy_(zh, non_blocking=True)
                    gd[b].copy_(gh, non_blocking=True)
                    ready[b].record(copy_stream)
                main.wait_event(ready[b])
                l, _ = step(zd[b], gd[b], out=grads[b])
                free[b].record(main)
                loss_host[b].copy_(l, non_blocking=True)
                done[b].record(main)
                if overlap:
                    if i >= 1:
                        done[b ^ 1].synchronize()            # host reads the previous step's losses
                        got.append(float(loss_host[b ^ 1][1]))
                else:
                    done[b].synchronize()
                    got.append(float(loss_host[b][1]))
            if overlap:
                done[(k - 1) % 2].synchronize()
                got.append(float(loss_host[(k - 1) % 2][1]))
            return got

        def time_e2e(masks_u8, overlap):
            st = e2e_state(masks_u8, overlap)
            run_e2e(3, masks_u8, overlap, st)
            barrier()
            ev0.record()
            got = run_e2e(e_steps, masks_u8, overlap, st)
            ev1.record()
            barrier()
            assert len(got) == e_steps and all(v == v for v in got)
            te = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e_ms = float(te.item()) / e_steps
            mask_bytes = 1 if masks_u8 else 4
            return {"value": pixels_per_step / (e_ms * 1e-3) / 1e9, "unit": "Gpixel/s",
                    "h2d_bytes_per_step": elems_per_gpu * (4 + mask_bytes) * world, "d2h_bytes_per_step": 7 * 4 * world,
                    "ms_per_step": e_ms, "steps": e_steps,
                    "how": ("fp32 logits + uint8 masks from pinned host memory, H2D of step i+1 on a copy stream under step i, "
                            "7 losses read back every step (one step late)") if overlap else
                           "fp32 logits + fp32 masks from pinned host memory, copy -> copy -> step -> read, no overlap"}

        e2e = time_e2e(True, True)
        e2e_variants["e2e_f32_masks_serial"] = time_e2e(False, False)

    # ---- parity of what was just timed (outside the timed region) ----------------------------------------------
    parity = composite_parity(step, dev_sets[0][0], dev_sets[0][1], weights, dev, world, rank)

    # ---- auxiliary line items (not the headline): the Dice-scoring kernel and the un-fused leaf path ----------
    aux = None
    if world == 1 and not args.no_aux:
        aux = measure_aux(dev)
    elif world > 1 and not args.no_aux and args.workload == "cfg2":
        # BASELINE configs[3] (cfg4): 432x3x512x512 over 8 GPUs = 54x3x512x512 per GPU.  Timed here on every rank with the
        # sharded step (max over ranks), next to the single-GPU step on the same shard on rank 0's GPU.
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
        barrier()
        sh = time_cfg4_shard(dev, step, peak)
        tt = torch.tensor([sh["us"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        barrier()
        np.random.seed(0)
        single = time_cfg4_shard(dev, fused.CompositeLossStep(weights, device=dev), peak) if rank == 0 else None
        barrier()
        if rank == 0:
            us_n = float(tt.item())
            aux = {"cfg4": {"what": f"BASELINE configs[3]: composite loss fwd+bwd, 54x3x512x512 f32 per GPU, global batch {54 * world}, "
                                    "sums all-reduced in-kernel; us per step = max over ranks",
                            "us_per_step": us_n, "gpixel_per_s": 54 * 512 * 512 * world / (us_n * 1e-6) / 1e9,
                            "single_gpu_same_shard_us": single["us"], "speedup_vs_1gpu": world * single["us"] / us_n,
                            "frac_of_hbm_peak_per_gpu": 12.0 * 54 * 3 * 512 * 512 / (us_n * 1e-6) / 1e9 / peak}}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg_bytes = 12.0 * elems_per_gpu  # read logits 4 + read masks 4 + write grad 4 per element, per GPU
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload if world == 1 else args.workload + "_sharded")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "kernel": "composite3_fused_v3_kernel" if (world == 1 or args.exchange == "p2p") else "composite3_stats_packed+allreduce+finalize+composite3_grad_v2",
                    "algorithmic_bytes_per_launch": alg_bytes}
        if args.workload == "cfg2" and (world == 1 or args.exchange == "p2p"):
            # what really bounds this kernel: its instruction stream.  Warp instructions per launch are a
            # property of the code and the shape (ncu smsp__inst_executed.sum of the committed capture); the issue peak is one
            # warp instruction per clock and SM sub-partition at the clock sampled during the timed region.
            try:
                ncu = json.load(open(os.path.join(ROOT, "profiles", "r2c_fused_v3_ncu_full.json")))["launches"][0]
                inst = float(ncu["smsp__inst_executed.sum"]["value"])
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                peak_issue = 4.0 * sms * mhz * 1e6
                roofline["instruction_stream"] = {
                    "warp_instructions_per_launch": inst, "achieved_ginst_per_s": inst / (ms_per_step * 1e-3) / 1e9,
                    "peak_ginst_per_s": peak_issue / 1e9, "frac": inst / (ms_per_step * 1e-3) / peak_issue,
                    "note": "informational: the step is bound by its issue + XU work (FFMA2 holds an issue slot for 2-3 cycles), "
                            "not by HBM; source of the count: profiles/r2c_fused_v3_ncu_full.json"}
            except Exception:
                pass
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            res = time_cpu_baseline(args.workload, budget_s=15.0)
            tb = min(res["times"])
            kind = reference_impl()[0]
            cpu_baseline = {"value": res["pixels"] / tb / 1e9, "unit": "Gpixel/s", "cores": res["cores"],
                            "kind": kind, "sample": f"{len(res['times'])} full steps of {args.workload} "
                            f"({res['shape'][0]}x{c}x{s}x{s}), best of; " + reference_kind_text(kind),
                            "ms_per_step": tb * 1e3}
        line = {
            "metric": "Gpixel/s fused loss fwd+bwd", "value": value, "unit": "Gpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, world, args.exchange),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps * world,
            "losses": [float(v) for v in losses.cpu()],
            "parity": parity,
            "e2e_variants": e2e_variants, "host": numa_note,
            "aux": aux,
        }
    else:
        line = None
    if world > 1:
        if hasattr(step, "close"):
            step.close()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
