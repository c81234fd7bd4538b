set -x
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo rc=$?
python bench.py --workload cfg2_5k_city --steps 200 --warmup 20 --graph > gpurun_out/r02_bench_cfg2_5k_city_graph.json 2>/dev/null; echo rc=$?
python bench.py --workload cfg3_100k_uniform2d --steps 100 --warmup 10 > gpurun_out/r02_bench_cfg3_100k_uniform2d.json 2>/dev/null; echo rc=$?
python bench.py --workload cfg1_1k_city --steps 200 --warmup 20 --graph > gpurun_out/r02_bench_cfg1_1k_city_graph.json 2>/dev/null; echo rc=$?
python tools/bench_index.py 1000000 8000000 32000000 > gpurun_out/r02_index_microbench.jsonl 2>/dev/null; cat gpurun_out/r02_index_microbench.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --verify-queries 0 --cpu-budget 1 > gpurun_out/r02_launches_run.json 2> /dev/null; echo rc=$?
ncu --set full --clock-control none --import-source on --launch-skip 19 -c 18 -o gpurun_out/r02_frame -f python tools/prof_one.py 1mc fused 2 > gpurun_out/r02_ncu_frame.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_reorder|k_pack_keys|k_cell_table" --launch-skip 24 -c 6 -o gpurun_out/r02_index -f python tools/bench_index.py 32000000 > gpurun_out/r02_ncu_index.log 2>&1; echo rc=$?
ls -la gpurun_out/r02_*
