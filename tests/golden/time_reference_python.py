"""Build-container only: time the REFERENCE's own Python code (under the shim of oracle/ref_shim.py)
on the configs[0]/configs[1] frames, single interpreter thread -- the like-for-like CPU number of
BASELINE.md section 3 item 2.  Writes profiles/r01_reference_python_cpu.json."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import ref_shim as S
from rcd_b200.host import workloads as W

out = {"host": "build container, 8-core Intel Xeon, Python 3.12 (one interpreter thread)", "runs": []}
for name, n, sample in (("cfg1_1k_city", 1000, 1000), ("cfg2_5k_city", 5000, 250)):
    frame = W.make_workload(name)
    f64 = W.frame_to_f64(frame)
    ref, index, det, vehicles = S.build_detector_A(f64)   # untimed warm build
    t0 = time.perf_counter()
    ref, index, det, vehicles = S.build_detector_A(f64)
    t_ingest = time.perf_counter() - t0
    ids = [f"v{i}" for i in range(0, n, max(1, n // sample))]
    t0 = time.perf_counter()
    for vid in ids:
        det.detect_collisions(vid)
    t_detect = (time.perf_counter() - t0) * n / len(ids)
    model = ref.collision_detection.CollisionPredictionModel(det)
    for v in vehicles:
        model.update_trajectory(v.id, v.position, 0.0)
        model.update_trajectory(v.id, ref.Position(v.position.x + v.velocity.x, v.position.y + v.velocity.y, v.position.z), 1.0)
    t0 = time.perf_counter()
    for vid in ids:
        model.predict_collisions(vid)
    t_predict = (time.perf_counter() - t0) * n / len(ids)
    frame_s = t_ingest + t_detect + t_predict
    out["runs"].append({"workload": name, "objects": n, "queried": len(ids), "ingest_s": t_ingest, "detect_s": t_detect,
                        "predict_s": t_predict, "frame_s": frame_s, "object_updates_per_s": n / frame_s})
    print(out["runs"][-1], flush=True)
json.dump(out, open("profiles/r01_reference_python_cpu.json", "w"), indent=1)
