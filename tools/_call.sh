timeout 600 python -m pytest tests/test_gpu_alerts.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
