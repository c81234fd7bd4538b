set -x
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 10 --warmup 3 --skip-configs4 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo rc=$?
tail -c 400 gpurun_out/r02_bench_n8.err
