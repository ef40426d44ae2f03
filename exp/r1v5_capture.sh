set -x
python bench.py > gpurun_out/bench_r1v5.json 2> gpurun_out/bench_r1v5.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1v5.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l5.log 2>&1
python bench.py --workload cfg5 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/cfg5_chk.json 2>/dev/null && \
ncu --set full --clock-control none --import-source on -k regex:dice_counts -c 2 -o gpurun_out/prof_dice_r1v5 -f python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_d5.log 2>&1
cat > /tmp/m.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from ecologysemanticsegmentation_b200 import ops
z = torch.randn(54, 3, 1024, 1024, device='cuda')
for _ in range(3):
    ops.masks_u8(z, None); ops.masks_u8(z, 0.8)
torch.cuda.synchronize()
PY
python /tmp/m.py && ncu --set full --clock-control none --import-source on -k regex:masks_u8 -c 2 -o gpurun_out/prof_masks_r1v5 -f python /tmp/m.py > gpurun_out/ncu_m5.log 2>&1
ls -la gpurun_out/*r1v5*
