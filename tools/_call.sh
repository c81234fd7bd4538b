timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02_pytest_gpu_final.log
python tools/ab_stage.py - | tee gpurun_out/r2x_ab.txt
python tools/bench_index.py 32000000 | tee -a gpurun_out/r2x_ab.txt
