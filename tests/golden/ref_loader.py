"""Load the UNMODIFIED reference loss path from /root/reference (this container only).

Test infrastructure.  Used solely by ``make_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, i.e. on
the GPU box).  Nothing in the product package imports this.

The reference package cannot be imported normally: ``ecology_semantic_segmentation/__init__.py``
eagerly pulls datasets -> albumentations (absent) and a private dataset.  We register a stub
parent package that carries only ``binary_cross_entropy = torch.nn.BCEWithLogitsLoss()``
(exactly ``ecology_semantic_segmentation/__init__.py:24``) and then exec the reference's own
``loss_functions.py`` / ``loss_composite.py`` files, unmodified, as its submodules.
``train_multiclass.losses_fn`` (train_multiclass.py:253-303) is pulled out of its file by AST,
because importing that module would drag in datasets and segmentation_models_pytorch.
"""
import ast
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ECO_REFERENCE_ROOT", "/root/reference")
PKG = "ecology_semantic_segmentation"
PKG_DIR = os.path.join(REF_ROOT, PKG)


def available() -> bool:
    return os.path.isfile(os.path.join(PKG_DIR, "loss_functions.py"))


def load():
    """Returns (loss_functions, loss_composite, train_multiclass_losses_fn)."""
    import numpy as np
    import torch

    if PKG not in sys.modules or not hasattr(sys.modules[PKG], "_eco_stub"):
        stub = types.ModuleType(PKG)
        stub.__path__ = [PKG_DIR]
        stub._eco_stub = True
        stub.binary_cross_entropy = torch.nn.BCEWithLogitsLoss()
        sys.modules[PKG] = stub

    mods = []
    for name in ("loss_functions", "loss_composite"):
        full = f"{PKG}.{name}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(PKG_DIR, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        mods.append(sys.modules[full])
    lf, lc = mods

    src = open(os.path.join(PKG_DIR, "train_multiclass.py")).read()
    tree = ast.parse(src)
    fn_node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "losses_fn")
    ns = {"np": np, "torch": torch}
    for k in ("cross_entropy_loss", "focal_loss", "classification_dice_loss", "cross_entropy_list",
              "binary_cross_entropy_list", "focal_list", "classification_dice_list", "dice_loss"):
        ns[k] = getattr(lf, k)
    exec(compile(ast.Module(body=[fn_node], type_ignores=[]), "train_multiclass.py[losses_fn]", "exec"), ns)
    return lf, lc, ns["losses_fn"]


def load_adjacent():
    """Functions next to the path, pulled out of their files by AST because importing
    those modules drags in datasets: (class-dim union from utils/subsets_union.py:8-32, batch-dim twin from
    train_multiclass.py:32-45, losses_fn of train_multiclass_sequential_densenetloss.py:272-362)."""
    import numpy as np
    import torch
    lf, lc, tm = load()

    def extract(rel, name, ns):
        path = os.path.join(PKG_DIR, rel)
        tree = ast.parse(open(path).read())
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
        exec(compile(ast.Module(body=[node], type_ignores=[]), rel + "[" + name + "]", "exec"), ns)
        return ns[name]

    base = {"np": np, "torch": torch}
    union_cls = extract(os.path.join("utils", "subsets_union.py"), "return_union_sets_descending_order", dict(base))
    union_bat = extract("train_multiclass.py", "return_union_sets_descending_order", dict(base))
    ns = dict(base)
    for k in ("cross_entropy_loss", "focal_loss", "classification_dice_loss", "cross_entropy_list",
              "binary_cross_entropy_list", "focal_list", "classification_dice_list", "dice_loss"):
        ns[k] = getattr(lf, k)
    seq = extract("train_multiclass_sequential_densenetloss.py", "losses_fn", ns)
    return union_cls, union_bat, seq
