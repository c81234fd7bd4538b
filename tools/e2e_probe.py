"""Scratch: where the end-to-end time of the summary delivery goes (1 M clustered frame, device-resident state)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine
frame, bounds = W.make_workload("cfg4_1m_clustered3d"), ((0, 0, 0), (31623, 31623, 100))
f2 = W.advance(frame, 0.05, np.random.default_rng(99), map_size=(31623.0, 31623.0))
n = len(frame["px"])
with FrameEngine(n, 32_000_000, world_bounds=bounds) as e:
    e.alerts_configure(int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000)
    ev = np.zeros(4 << 20, N.ALERT_EVENT_DTYPE); rc = np.zeros(n, np.uint32)
    pat = np.full(n, 2, np.uint8)
    def run(kind, reps=8):
        ts = []
        for r in range(reps):
            e.upload(frame if r % 2 == 0 else f2); e.set_patterns(pat); e.sync()
            t0 = time.perf_counter()
            e.step(N.MODE_PREDICT, with_detect=True)
            if kind == "frame":
                e.sync()
            elif kind == "summary":
                e.summary_begin(1000.0 + r); got = e.summary_finish(ev, rc)
            elif kind == "alerts":
                got = e.alerts_update(1000.0 + r, cap=0)
            ts.append((time.perf_counter() - t0) * 1e3)
        return round(float(np.median(ts[2:])), 3), (got[1] if kind != "frame" else None)
    for kind in ("frame", "summary", "frame", "alerts", "summary"):
        print(kind, run(kind), flush=True)
