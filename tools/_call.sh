for v in old new4 rq0 pk0; do
  echo "== $v"; RCD_B200_LIB=$PWD/build/ab/lib_$v.so timeout 120 python tools/bench_index.py 1000000 2000000 2>&1 | tail -2
done | tee gpurun_out/r2o_index_variants.txt
python tools/ab_stage.py build/ab/lib_old.so build/ab/lib_new4.so build/ab/lib_old.so build/ab/lib_new4.so | tee -a gpurun_out/r2o_index_variants.txt
