python tools/ab_stage.py build/ab/lib_seg.so build/ab/lib_res.so build/ab/lib_res_r5.so build/ab/lib_res_r3.so build/ab/lib_seg.so build/ab/lib_res.so | tee gpurun_out/r2t_ab.txt
RCD_B200_LIB=$PWD/build/ab/lib_res.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
