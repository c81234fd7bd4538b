set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --workload cfg5_10m_skew3d --objects-per-gpu 312500 --verify-kind counts --verify-queries 400 --no-extras > gpurun_out/r02_bench_cfg5_counts_n2.json 2> gpurun_out/r02_bench_cfg5_counts_n2.err; echo rc=$?
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_multi_gpu_check_n2_tail.log
tail -c 600 gpurun_out/r02_bench_n2.err; tail -c 600 gpurun_out/r02_bench_cfg5_counts_n2.err
