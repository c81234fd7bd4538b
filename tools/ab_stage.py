"""Scratch: stage times of the fused bench frame (1 M clustered) for the library in RCD_B200_LIB (A/B kernel variants).
usage: python tools/ab_stage.py [lib.so ...]   -- without arguments: the in-tree library, in this process."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    for lib in sys.argv[1:]:
        env = dict(os.environ)
        if lib != "-":
            env["RCD_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        print(lib, r.stdout.strip() or r.stderr[-400:], flush=True)
    sys.exit(0)
import numpy as np
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine
frame, bounds = W.make_workload("cfg4_1m_clustered3d"), ((0, 0, 0), (31623, 31623, 100))
n = len(frame["px"])
with FrameEngine(n, 32_000_000, world_bounds=bounds, profile=True) as e:
    e.upload(frame)
    e.set_patterns(np.full(n, 2, np.uint8))
    best = None
    for r in range(6):
        e.invalidate()
        e.step(N.MODE_PREDICT, with_detect=True)
        e.sync()
        ms = e.stage_ms(N.MODE_PREDICT)
        if best is None or ms["total"] < best["total"]:
            best = ms
    print(json.dumps({k: round(v, 3) for k, v in best.items() if v}))
