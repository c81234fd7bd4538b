for v in ag1 ag16 ag32 ag64; do echo "== $v"; RCD_B200_LIB=$PWD/build/ab/lib_$v.so timeout 300 python tools/bench_alerts.py 2>&1 | head -5 | cut -c1-330; done | tee gpurun_out/r2u_alerts.txt
RCD_B200_LIB=$PWD/build/ab/lib_ag32.so timeout 600 python -m pytest tests/test_gpu_alerts.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
for v in ag1 ag32; do echo "== $v"; RCD_B200_LIB=$PWD/build/ab/lib_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --verify-queries 0 --cpu-budget 1 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin); e=d['e2e']; print(d['ms_per_step'], e['ms_per_step'], e['value'], e['paced_arrivals']['p99_ms'], e['other_inflight'])"; done | tee -a gpurun_out/r2u_alerts.txt
