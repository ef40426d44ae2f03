"""bench.py's reference arm runs without a GPU (it times the CPU oracle port) and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_cfg5():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg5",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gpixel/s" and d["higher_is_better"] is True
    assert d["metric"] == "Gpixel/s Dice eval" and d["value"] > 0 and d["ms_per_step"] > 0
    from oracle import make_ref
    kind = "reference" if make_ref.staged() else "port"   # the unmodified reference when oracle/_ref is staged
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "sample" in d["cpu_baseline"]
    # both arms print the same `config` object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict("cfg5", 1)


def test_make_ref_stages_the_unmodified_reference():
    """oracle/make_ref.py: byte copies of the reference's two files (when /root/reference exists), loadable, and equal to
    the port on a small case."""
    import numpy as np
    import pytest
    import torch
    from oracle import make_ref, torch_port as tp
    if not make_ref.stage():
        pytest.skip("neither /root/reference nor a staged oracle/_ref here")
    lf, lc = make_ref.load()
    torch.manual_seed(0)
    p = torch.sigmoid(torch.randn(2, 3, 12, 12))
    g = (torch.rand(2, 3, 12, 12) > 0.5).float()
    np.random.seed(1)
    a = [float(v) for v in lc.losses_fn(p, g, True)]
    np.random.seed(1)
    b = [float(v) for v in tp.losses_composite(p, g, True)]
    assert a == b


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
