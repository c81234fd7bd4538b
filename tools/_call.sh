timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --verify-queries 500 --cpu-budget 1 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin); e=d['e2e']; print('default', d['ms_per_step'], d['latency_ms']['p99'], 'e2e', e['ms_per_step'], e['value'], e['paced_arrivals']['p99_ms'], e['other_inflight'], d['verify']['ok'])"
python bench.py --steps 20 --warmup 3 --verify-queries 500 --cpu-budget 1 --graph 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin); e=d['e2e']; print('graph', d['ms_per_step'], d['latency_ms']['p99'], 'e2e', e['ms_per_step'], e['value'], d['graph_replays'], d['verify']['ok'])"
