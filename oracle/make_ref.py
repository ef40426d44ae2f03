"""Stage the UNMODIFIED reference loss path under ``oracle/_ref/`` so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``oracle/_ref/`` is
git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored, so the files staged
here ride along with the snapshot exactly like a built ``.so``.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg then time the reference's own ``loss_composite.losses_fn`` / ``loss_functions.dice_loss``
(``kind: "reference"``) and fall back to the op-for-op port ``oracle/torch_port.py`` (``kind: "port"``) only when
the staging is absent.  ``tests/test_oracle_vs_reference.py`` checks the port against the same files bit for bit.

    python -m oracle.make_ref            # stage (no-op when /root/reference is absent)

The reference is pure Python: there is nothing to compile.  The staged files are byte copies; a manifest with their
sha256 is written next to them and checked at load time.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ECO_REFERENCE_ROOT", "/root/reference")
PKG = "ecology_semantic_segmentation"
DEST = os.path.join(HERE, "_ref", PKG)
FILES = ("loss_functions.py", "loss_composite.py")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose=False) -> bool:
    """Copy the two files of the path from the reference checkout.  Returns True when oracle/_ref is usable."""
    src_dir = os.path.join(REF_ROOT, PKG)
    if not all(os.path.isfile(os.path.join(src_dir, f)) for f in FILES):
        return staged()
    os.makedirs(DEST, exist_ok=True)
    man = {"source": src_dir, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(DEST, f))
        man["files"][f] = _sha(os.path.join(DEST, f))
    with open(MANIFEST, "w") as fh:
        json.dump(man, fh, indent=1)
    if verbose:
        print("staged", ", ".join(FILES), "->", DEST)
    return True


def staged() -> bool:
    if not os.path.isfile(MANIFEST):
        return False
    try:
        man = json.load(open(MANIFEST))
        return all(_sha(os.path.join(DEST, f)) == h for f, h in man["files"].items())
    except Exception:
        return False


def load():
    """(loss_functions, loss_composite) modules of the staged, unmodified reference, imported under a stub parent
    package that carries only ``binary_cross_entropy = torch.nn.BCEWithLogitsLoss()`` (ess/__init__.py:24; the real
    ``__init__`` drags in datasets and albumentations).  Raises if nothing is staged."""
    import torch
    if not staged():
        raise RuntimeError("oracle/_ref is not staged (run `python -m oracle.make_ref` where /root/reference exists)")
    name = "_eco_ref_pkg." + PKG   # private parent so that it never collides with tests/golden/ref_loader.py's stub
    if name not in sys.modules:
        top = types.ModuleType("_eco_ref_pkg")
        top.__path__ = []
        sys.modules.setdefault("_eco_ref_pkg", top)
        stub = types.ModuleType(name)
        stub.__path__ = [DEST]
        stub.binary_cross_entropy = torch.nn.BCEWithLogitsLoss()
        sys.modules[name] = stub
    mods = []
    for f in FILES:
        full = name + "." + f[:-3]
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(DEST, f))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        mods.append(sys.modules[full])
    return tuple(mods)


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref staged:", ok)
