"""Class-level end-to-end benchmark: the reference perf harness's frame through the DROP-IN CLASSES
(rcd_b200.host: SpatialIndex / CollisionDetector / CollisionPredictionModel / AlertManager), one Python call per vehicle
like the reference's callers, on BASELINE configs[0] (1000 vehicles) and configs[1] (5000 vehicles):

  harness frame   performance_test.py:794-813 -- update every vehicle, detect_collisions(id) for every vehicle,
                  predict_collisions(id) for every vehicle
  warning frame   warning_system.py:638-714   -- update_vehicle + update_trajectory per message, then
                  predict_collisions(id) + process_collision_risks(risks) for every vehicle (_detect_all_vehicles)

    python tools/bench_classes.py [--frames K]
Prints one JSON line per (config, loop) with frames/s next to the reference's published numbers."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rcd_b200.host import workloads as W  # noqa: E402
from rcd_b200.host.collision_detection import CollisionDetector, CollisionPredictionModel  # noqa: E402
from rcd_b200.host.models import Position, Vector, Vehicle  # noqa: E402
from rcd_b200.host.spatial_index import SpatialIndex  # noqa: E402
from rcd_b200.host.warning_system import AlertManager  # noqa: E402

PUBLISHED = {1000: {"frames_per_s": 9.44, "p50_ms": 67.8, "p99_ms": 314.57,
                    "source": "results/optimized_perf_test_1000vehicles_1000tps_30s_20250315_034842_summary.txt:5-21"},
             5000: {"frames_per_s": 1 / 73.30, "p50_ms": 73275.0, "p99_ms": 73275.0,
                    "source": "results/perf_test_5000vehicles_5000tps_60s_20250315_034432_summary.txt:5-21"}}
TYPES = ["car", "truck", "bus", "motorcycle", "drone"]


def vehicles_of(frame, ts):
    f = {k: v.tolist() for k, v in frame.items()}
    return [Vehicle(id=f"vehicle-{i}", position=Position(f["px"][i], f["py"][i], f["pz"][i]),
                    velocity=Vector(f["vx"][i], f["vy"][i], f["vz"][i]), acceleration=Vector(f["ax"][i], f["ay"][i], f["az"][i]),
                    heading=f["heading"][i], size=f["size"][i], type=TYPES[f["type"][i]], timestamp=ts) for i in range(len(f["px"]))]


def run(n, loop, n_frames, warmup=3):
    frames = [W.reference_city_frame(n, 1234 if n == 1000 else 1235)]
    rng = np.random.default_rng(99)
    for _ in range(1, n_frames + warmup):  # the generator moves every vehicle between frames (performance_test.py:147-195)
        frames.append(W.advance(frames[-1], 0.1, rng))
    index = SpatialIndex()
    det = CollisionDetector(index)
    model = CollisionPredictionModel(det)
    alerts = AlertManager() if loop == "warning" else None
    lat, n_risks, n_alerts = [], 0, 0
    t_gen = 0.0
    for k, frame in enumerate(frames):
        ts = 1000.0 + 0.1 * k
        t0 = time.perf_counter()
        vehicles = vehicles_of(frame, ts)  # (the reference's generator hands out Vehicle objects: not part of its timed region)
        t1 = time.perf_counter()
        t_gen += t1 - t0
        for v in vehicles:
            det.update_vehicle(v)
            model.update_trajectory(v.id, v.position, v.timestamp)
        risks_seen = 0
        if loop == "harness":
            for v in vehicles:
                risks_seen += len(det.detect_collisions(v.id))
            for v in vehicles:
                risks_seen += len(model.predict_collisions(v.id))
        else:
            for vid in list(det.vehicle_cache.keys()):
                risks = model.predict_collisions(vid)
                if risks:
                    risks_seen += len(risks)
                    n_alerts += len(alerts.process_collision_risks(risks))
        dt = time.perf_counter() - t1
        if k >= warmup:
            lat.append(dt * 1e3)
            n_risks += risks_seen
    lat = np.array(lat)
    pub = PUBLISHED[n]
    return {"config": f"configs[{0 if n == 1000 else 1}]: reference perf-test generator, {n} vehicles, 10 km map", "loop": loop,
            "api": "per-vehicle calls on the drop-in classes (rcd_b200.host.*), results served from one GPU frame per mode",
            "frames": len(lat), "frames_per_s": 1e3 / float(lat.mean()), "object_updates_per_s": n * 1e3 / float(lat.mean()),
            "latency_ms": {"mean": float(lat.mean()), "p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "max": float(lat.max())},
            "risks_per_frame": n_risks / len(lat), "alerts_returned": n_alerts,
            "reference_published": pub, "speedup_vs_published_frames_per_s": (1e3 / float(lat.mean())) / pub["frames_per_s"]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    a = ap.parse_args()
    for n in (1000, 5000):
        for loop in ("harness", "warning"):
            print(json.dumps(run(n, loop, a.frames)), flush=True)
