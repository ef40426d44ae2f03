"""In-tree build of ``libecoloss.so`` (hand-written sm_100a kernels + the C ABI of include/ecoloss.h).

    python -m ecologysemanticsegmentation_b200.csrc.build [--force]

nvcc cross-compiles without a GPU.  The .so stays in this directory so that it travels with the
repo snapshot to the GPU box and shows up as a loaded in-tree library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["eco_api.cu", "eco_leaf.cu", "eco_composite.cu", "eco_eval.cu", "eco_masks.cu", "eco_softce.cu", "eco_union.cu", "eco_frames.cu"]
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith(".cuh")) + [os.path.join(ROOT, "include", "ecoloss.h")]
LIB = os.path.join(HERE, "libecoloss.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",   # host code restates third-party double arithmetic operation by operation
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    extra = os.environ.get("ECO_EXTRA_NVCC_FLAGS", "").split()   # experiment switches (-DECO_...), see DESIGN.md section 4
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{log}")
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            logs = list(ex.map(compile_one, jobs))
        if verbose:
            for l in logs:
                print(l)
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
