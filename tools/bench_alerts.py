"""Alert-table microbenchmark: fold the pairs of a full bench frame (configs[3], 1 M
objects, detect + predict) into the device table; first frame (every alert is created, events come back) and
steady state (alerts refreshed, nothing to report), next to the reference's dict walk (oracle port) on a sample.
    python tools/bench_alerts.py
Prints JSON lines."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rcd_b200.host import _native as N, workloads as W  # noqa: E402
from rcd_b200.host.engine import FrameEngine  # noqa: E402

frame = W.make_workload("cfg4_1m_clustered3d")
n = len(frame["px"])
with FrameEngine(n, 32_000_000, world_bounds=((0, 0, 0), (31623, 31623, 100))) as e:
    e.upload(frame)
    e.set_patterns(np.full(n, 2, np.uint8))
    e.step(N.MODE_DETECT)
    e.step(N.MODE_PREDICT, append=True)
    c = e.counts()
    e.alerts_configure(16_000_000)
    rows = []
    for it, now in enumerate((10.0, 10.5, 11.0, 11.5)):
        e.sync()
        t0 = time.perf_counter()
        ev, st = e.alerts_update(now, cap=0 if it else None)  # later frames: counters only
        dt = time.perf_counter() - t0
        rows.append({"what": "rcd_alerts_update, first frame (all created, events copied to the host)" if it == 0
                     else "rcd_alerts_update, steady state", "pairs": c["n_pairs"], "alerting": st["n_created"] + st["n_changed"] + st["n_refreshed"],
                     "created": st["n_created"], "changed": st["n_changed"], "events_to_host": len(ev), "ms": round(dt * 1e3, 3),
                     "risks_per_s": round((st["n_created"] + st["n_changed"] + st["n_refreshed"]) / dt)})
    t0 = time.perf_counter()
    ev, st = e.alerts_expire(60.0)
    rows.append({"what": "rcd_alerts_expire (rebuild, all expired)", "expired": st["n_expired"], "ms": round((time.perf_counter() - t0) * 1e3, 3)})
    pairs = e.download(sort=False)
for r in rows:
    print(json.dumps(r))
# the reference's dict walk (warning_system.py:259-285, 120-197) on a sample of the same risks
def priority(risk, ttc):
    if risk >= 0.8 and ttc < 3.0:
        return 3
    if risk >= 0.8 or ttc < 5.0:
        return 2
    return 1 if risk >= 0.6 else 0


def dict_walk(table, risks, now):
    for i, j, risk, ttc in risks:
        if risk < 0.3:
            continue
        a = table.get((i, j))
        if a is None:
            table[(i, j)] = [risk, ttc, priority(risk, ttc), now]
        else:
            a[0], a[1], a[2], a[3] = risk, ttc, priority(risk, ttc), now


sample = pairs[pairs["priority"] >= 0][:300_000]
risks = [(int(p["i"]), int(p["j"]), float(p["risk"]), float(p["ttc"])) for p in sample]
table = {}
t0 = time.perf_counter()
dict_walk(table, risks, 10.0)
dict_walk(table, risks, 10.5)
dt = time.perf_counter() - t0
print(json.dumps({"what": "reference dict walk (Python, 1 core), create + refresh", "risks": 2 * len(risks), "risks_per_s": round(2 * len(risks) / dt)}))
