set -x
python bench.py --steps 5 --warmup 3 --verify-kind counts --verify-queries 1000 > gpurun_out/r2l_counts.json 2> gpurun_out/r2l_counts.err; echo rc=$?
for v in old base lb16 lb32 vl vl4 c4 i12c4 i12c5vl; do
  lib=build/ab/lib_$v.so; [ $v = base ] && lib=realtime-collision-detection_b200/librcd_b200.so
  echo "== $v"; RCD_B200_LIB=$PWD/$lib python tools/bench_index.py 8000000 32000000 2>&1 | tail -2
done | tee gpurun_out/r2l_index_variants.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_onesweep_pass|k_reorder|k_pack_keys" --launch-skip 20 -c 5 -o gpurun_out/r2l_index python tools/bench_index.py 32000000 > gpurun_out/r2l_ncu.log 2>&1; echo ncu rc=$?
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2l_pytest.log
