"""Scratch: per-stage device times of a few workloads (run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from rcd_b200.host import workloads as W, _native as N
from rcd_b200.host.engine import FrameEngine

def run(name, frame, bounds, modes=("detect", "predict"), reps=5, max_pairs=8_000_000):
    n = len(frame["px"])
    with FrameEngine(n, max_pairs, world_bounds=bounds, profile=True) as e:
        e.upload(frame)
        e.set_patterns(np.full(n, 2, np.uint8))
        for mode in modes:
            m = N.MODE_DETECT if mode == "detect" else N.MODE_PREDICT
            best = None
            for r in range(reps):
                e.invalidate()
                t0 = time.perf_counter()
                e.step(m)
                e.sync()
                wall = (time.perf_counter() - t0) * 1e3
                ms = e.stage_ms(m)
                if best is None or ms["total"] < best["total"]:
                    best = dict(ms, wall=wall)
            c = e.counts()
            print(json.dumps({"workload": name, "n": n, "mode": mode, **{k + "_ms": round(v, 4) for k, v in best.items()},
                              "cand": c["n_candidates"], "pairs": c["n_pairs"], "exact": c["n_exact"]}), flush=True)

which = sys.argv[1:] or ["5k", "100k", "1m", "skew"]
if "5k" in which:
    run("cfg2_5k_city", W.make_workload("cfg2_5k_city"), ((0, 0, 0), (10000, 10000, 0)))
if "100k" in which:
    run("cfg3_100k", W.make_workload("cfg3_100k_uniform2d"), ((0, 0, 0), (10000, 10000, 0)))
if "1m" in which:
    run("cfg4_1m", W.make_workload("cfg4_1m_clustered3d"), ((0, 0, 0), (31623, 31623, 100)), max_pairs=50_000_000)
if "1m_uniform" in which:
    run("1m_uniform2d", W.uniform_frame(1_000_000, 5, map_size=31623.0), ((0, 0, 0), (31623, 31623, 0)))
if "skew" in which:
    f = W.hotspot_frame(1_250_000, 2003, 35355.0, 25, zipf_s=1.0)
    run("skew_1.25m", f, ((0, 0, 0), (35355, 35355, 100)), max_pairs=50_000_000)
if "skew_uni" in which:
    f = W.hotspot_frame(1_250_000, 2003, 35355.0, 25, zipf_s=1.0, radial_law="uniform")
    run("skew_1.25m_uniformdisc", f, ((0, 0, 0), (35355, 35355, 100)), max_pairs=50_000_000)
