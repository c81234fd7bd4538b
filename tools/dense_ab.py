"""Scratch: stage times of detect / predict on a very dense single hotspot (A/B between library builds via RCD_B200_LIB)."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
frame = W.hotspot_frame(n, 5, 4000.0, 1, radius_range=(1500.0, 1500.0))
with FrameEngine(n, 60_000_000, world_bounds=((0, 0, 0), (4000, 4000, 100)), profile=True) as e:
    e.upload(frame)
    e.set_patterns(np.full(n, 2, np.uint8))
    for mode, name in ((N.MODE_DETECT, "detect"), (N.MODE_PREDICT, "predict"), (N.MODE_PREDICT, "fused")):
        if name == "fused" and not hasattr(N.load(), "rcd_apply_records"):
            continue
        for r in range(2):
            e.invalidate()
            e.step(mode, with_detect=True) if name == "fused" else e.step(mode)
            e.sync()
        ms = e.stage_ms(mode); c = e.counts()
        print(json.dumps({"mode": name, "pairs": c["n_pairs"], "cand": c["n_candidates"], **{k: round(v, 3) for k, v in ms.items() if v > 0}}))
