set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_n2_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_n2_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?
