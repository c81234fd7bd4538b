set -x
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1gpu_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_1gpu_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
python tools/bench_alerts.py > gpurun_out/r02_alerts_microbench.jsonl 2>&1; head -4 gpurun_out/r02_alerts_microbench.jsonl | cut -c1-200
