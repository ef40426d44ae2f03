set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python profiles/kernel_rooflines.py > gpurun_out/kernel_rooflines_r1v5.jsonl 2>/dev/null; tail -4 gpurun_out/kernel_rooflines_r1v5.jsonl
