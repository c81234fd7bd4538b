for v in al_old al8 al12; do echo "== $v"; RCD_B200_LIB=$PWD/build/ab/lib_$v.so timeout 300 python tools/bench_alerts.py 2>&1 | sed -n 2,4p | cut -c1-200
RCD_B200_LIB=$PWD/build/ab/lib_$v.so timeout 300 python bench.py --steps 20 --warmup 3 --verify-queries 0 --cpu-budget 1 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin); e=d['e2e']; print(d['ms_per_step'], 'e2e', e['ms_per_step'], e['value'], e['other_inflight'])"; done | tee gpurun_out/r2y_alerts.txt
RCD_B200_LIB=$PWD/build/ab/lib_al12.so timeout 600 python -m pytest tests/test_gpu_alerts.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
