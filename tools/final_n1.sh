set -x
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo rc=$?
python bench.py --workload cfg2_5k_city --steps 200 --warmup 20 --graph > gpurun_out/r02_bench_cfg2_5k_city_graph.json 2>/dev/null; echo rc=$?
python bench.py --workload cfg3_100k_uniform2d --steps 100 --warmup 10 > gpurun_out/r02_bench_cfg3_100k_uniform2d.json 2>/dev/null; echo rc=$?
python bench.py --workload cfg1_1k_city --steps 200 --warmup 20 --graph > gpurun_out/r02_bench_cfg1_1k_city_graph.json 2>/dev/null; echo rc=$?
python tools/bench_classes.py --frames 20 > gpurun_out/r02_class_level.jsonl 2>&1; echo rc=$?
python tools/bench_alerts.py > gpurun_out/r02_alerts_microbench.jsonl 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --verify-queries 0 --cpu-budget 1 > gpurun_out/r02_launches_run.json 2> /dev/null; echo rc=$?
ncu --set full --clock-control none --import-source on --launch-skip 18 -c 17 -o gpurun_out/r02_frame -f python tools/prof_one.py 1mc fused 2 > gpurun_out/r02_ncu_frame.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:k_alert_update --launch-skip 4 -c 2 -o gpurun_out/r02_alerts -f python tools/bench_alerts.py > gpurun_out/r02_ncu_alerts.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:k_apply_records -c 2 -o gpurun_out/r02_ingest -f python tools/bench_ingest.py 1000000 > gpurun_out/r02_ncu_ingest.log 2>&1; echo rc=$?
ls -la gpurun_out/r02_*
