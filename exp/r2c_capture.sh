set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
python profiles/kernel_rooflines.py > gpurun_out/r2c_kernel_rooflines.jsonl 2>/dev/null; tail -12 gpurun_out/r2c_kernel_rooflines.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:composite3_fused_v3 -s 3 -c 2 -o gpurun_out/r2c_prof_fused -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-aux > gpurun_out/r2c_ncu_f.log 2>&1
ls -la gpurun_out/r2c_*
