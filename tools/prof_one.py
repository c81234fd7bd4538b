"""Scratch: run one workload / mode a few times (for ncu)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine
name, mode, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
if name == "100k":
    frame, bounds = W.make_workload("cfg3_100k_uniform2d"), ((0, 0, 0), (10000, 10000, 0))
elif name == "1m":
    frame, bounds = W.uniform_frame(1_000_000, 5, map_size=31623.0), ((0, 0, 0), (31623, 31623, 0))
elif name == "1mc":
    frame, bounds = W.make_workload("cfg4_1m_clustered3d"), ((0, 0, 0), (31623, 31623, 100))
elif name == "cfg5":
    side = 100000.0 * (1_250_000 / 10_000_000) ** 0.5
    frame, bounds = W.hotspot_frame(1_250_000, 2003, side, 25, zipf_s=1.0), ((0, 0, 0), (side, side, 100))
n = len(frame["px"])
with FrameEngine(n, 256_000_000 if name == "cfg5" else 40_000_000, world_bounds=bounds) as e:
    e.upload(frame)
    e.set_patterns(np.full(n, 2, np.uint8))
    for r in range(reps):
        e.invalidate()
        if mode == "fused":
            e.step(N.MODE_PREDICT, with_detect=True)
        else:
            e.step(N.MODE_DETECT if mode == "detect" else N.MODE_PREDICT)
        e.sync()
    print(e.counts())
