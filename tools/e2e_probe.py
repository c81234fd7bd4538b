"""Scratch: where the end-to-end time goes (1 M clustered frame): pipelined loop of bench.py with pieces switched off."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine, FRAME_FIELDS
frame, bounds = W.make_workload("cfg4_1m_clustered3d"), ((0, 0, 0), (31623, 31623, 100))
f2 = W.advance(frame, 0.05, np.random.default_rng(99), map_size=(31623.0, 31623.0))
n = len(frame["px"])
pin, dev = [], []
for f in (frame, f2):
    p = {k: torch.from_numpy(f[k]).pin_memory() for k in FRAME_FIELDS}
    p["type"] = torch.from_numpy(f["type"]).pin_memory()
    p["pattern"] = torch.full((n,), 2, dtype=torch.uint8).pin_memory()
    pin.append(p)
    dev.append({k: v.cuda() for k, v in p.items()})
with FrameEngine(n, 32_000_000, world_bounds=bounds) as e:
    e.alerts_configure(25_000_000)
    ev = torch.empty((4 << 20) * 40, dtype=torch.uint8).pin_memory().numpy().view(N.ALERT_EVENT_DTYPE)
    rc = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    def submit(k, host):
        src = (pin if host else dev)[k % 2]
        (e.upload_host_ptrs if host else e.upload_device)(n, [src[f].data_ptr() for f in FRAME_FIELDS], src["type"].data_ptr(), 0)
        (e.set_patterns_host_ptr if host else e.set_patterns_device)(n, src["pattern"].data_ptr())
        e.step(N.MODE_PREDICT, with_detect=True)
    def run(host, delivery, steps=12):
        now = 1000.0
        def begin():
            nonlocal now
            now += 0.05
            if delivery == "summary": e.summary_begin(now)
        def finish():
            if delivery == "summary": e.summary_finish(ev, rc)
        submit(0, host); begin()
        t0 = None
        for k in range(1, steps + 1):
            if k == 3:
                e.sync(); t0 = time.perf_counter(); k0 = k
            if k < steps: submit(k, host)
            finish()
            if k < steps: begin()
        e.sync()
        return round((time.perf_counter() - t0) / (steps - k0 + 0) * 1e3, 3)
    for host in (False, True):
        for delivery in ("none", "summary"):
            run(host, delivery, 6)
            print("host" if host else "device", delivery, run(host, delivery, 14), "ms/frame", flush=True)
