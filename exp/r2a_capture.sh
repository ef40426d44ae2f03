set -x
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-aux > gpurun_out/bench_r2a_chk.json 2> gpurun_out/bench_r2a_chk.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-aux > gpurun_out/ncu_l_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:composite3_fused_v3 -s 3 -c 2 -o gpurun_out/prof_fused_r2a -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-aux > gpurun_out/ncu_f_r2a.log 2>&1
ls -la gpurun_out/*r2a*
