"""Generate tests/golden/*.npz by running the REFERENCE's own bytecode (under the shim of
oracle/ref_shim.py, repairs R1-R5 of SURVEY.md 8c) on seeded synthetic frames.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim as S  # noqa: E402
from rcd_b200.host import workloads as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _frame_arrays(f):
    return {f"frame_{k}": v for k, v in f.items()}


def _risk_table(risks, ncol=9):
    if not risks:
        return np.zeros((0, ncol))
    return np.array([[float(x) for x in r[:ncol]] for r in risks], dtype=np.float64)


def detect_fixture(name, frame, R=100.0, T=10.0):
    f64 = W.frame_to_f64(frame)
    ref = S.run_detect_A(f64, R, T)
    np.savez_compressed(
        os.path.join(HERE, name), **_frame_arrays(frame), R=R, T=T,
        candidates=np.array(ref["candidates"], np.int32).reshape(-1, 2),
        potentials=np.array([[p[0], p[1], p[2], p[3]] for p in ref["potentials"]], np.float64).reshape(-1, 4),
        risks=_risk_table(ref["risks"]),
        stat_potential=ref["stats"]["potential_collisions"],
        stat_high=ref["stats"]["high_risk_collisions"])
    print(name, len(ref["candidates"]), len(ref["potentials"]), len(ref["risks"]))


def predict_fixture(name, frame, pattern):
    f64 = W.frame_to_f64(frame)
    ref = S.run_predict_A(f64, pattern)
    np.savez_compressed(os.path.join(HERE, name), **_frame_arrays(frame), pattern=pattern,
                        risks=_risk_table(ref["risks"]),
                        is_predicted=np.array([r[9] for r in ref["risks"]], bool))
    print(name, len(ref["risks"]))


def implB_fixture(name, frame, has_history):
    f64 = W.frame_to_f64(frame)
    ref = S.run_B(f64, has_history)
    np.savez_compressed(os.path.join(HERE, name), **_frame_arrays(frame), has_history=has_history,
                        candidates=np.array(ref["candidates"], np.int32).reshape(-1, 2),
                        risks=_risk_table(ref["risks"], 8))
    print(name, len(ref["candidates"]), len(ref["risks"]))


def scalars_fixture(name):
    rng = np.random.default_rng(99)
    # trajectory patterns (collision_detection.py:623-711)
    hist_flat, hist_off, classes = [], [0], []
    hs = []
    for _k in range(400):
        n = int(rng.integers(0, 14))
        mode = int(rng.integers(0, 5))
        t = np.cumsum(rng.choice([0.0, 0.5, 1.0], n)) if mode != 3 else rng.uniform(0, 5, n)
        p0 = rng.uniform(0, 100, 3)
        v = rng.uniform(-3, 3, 3) * (mode > 0) * rng.choice([0.01, 1])
        a = rng.uniform(-1, 1, 3) * (mode in (2, 4)) * rng.choice([0.05, 1])
        h = [tuple(p0 + v * tt + 0.5 * a * tt * tt) + (float(tt),) for tt in t]
        hs.append(h)
        hist_flat.extend(h)
        hist_off.append(len(hist_flat))
    classes = S.run_pattern_A(hs)
    # alert gate + priority (warning_system.py:259-311)
    rt = [(float(r), float(t)) for r in (0.0, 0.1, 0.29999, 0.3, 0.45, 0.59999, 0.6, 0.79999, 0.8, 0.95, 1.0)
          for t in (0.0, 0.1, 2.9, 3.0, 4.9, 5.0, 9.0, 19.4)]
    prio = S.run_priority_A(rt)
    # grid ids (spatial_index.py:97-112), incl. negative coordinates (trunc toward zero)
    pts = rng.uniform(-3000, 3000, (200, 3))
    pts[:5] = [(-0.5, -999.9, -0.1), (1500.2, -2500.0, 250.0), (999.999, 1000.0, 99.9), (0, 0, 0),
               (-1000.0, -1000.0, -100.0)]
    gids = {lvl: np.array(S.run_grid_id_A([tuple(p) for p in pts], lvl), np.int64) for lvl in range(4)}
    # radius queries (spatial_index.py:229-271)
    frame = W.uniform_frame(800, 5, map_size=700.0, drone_fraction=0.4)
    f64 = W.frame_to_f64(frame)
    q = np.concatenate([rng.uniform(-60, 760, (60, 3)) * [1, 1, 0.15],
                        np.stack([f64["px"][:20], f64["py"][:20], f64["pz"][:20]], 1)])
    q = q.astype(np.float32).astype(np.float64)
    near, noff = [], [0]
    for ids in S.run_nearby_A(f64, [tuple(x) for x in q], 90.0):
        near.extend(ids)
        noff.append(len(near))
    np.savez_compressed(
        os.path.join(HERE, name),
        hist=np.array(hist_flat, np.float64).reshape(-1, 4), hist_off=np.array(hist_off, np.int64),
        hist_class=np.array(classes, np.int32),
        prio_in=np.array(rt, np.float64), prio_out=np.array(prio, np.int32),
        grid_pts=pts, grid_l0=gids[0], grid_l1=gids[1], grid_l2=gids[2], grid_l3=gids[3],
        **{f"near_{k}": v for k, v in _frame_arrays(frame).items()},
        near_q=q, near_radius=90.0, near_ids=np.array(near, np.int32), near_off=np.array(noff, np.int64))
    print(name, np.bincount(classes), len(near))


def ingest_fixture(name):
    """EarlyWarningSystem._handle_vehicle_position (warning_system.py:638-678, shim repair R7) on the
    message texts of tests/ingest_cases.py: per message, dropped or the Vehicle fields."""
    import json
    from tests import ingest_cases as C
    texts = C.seeded_messages(400, 77) + [t for t, _ in C.edge_messages()]
    got = S.run_handle_position_A(texts)
    rows = []
    for t, g in zip(texts, got):
        rows.append({"text": t, "vehicle": None if g is None else [g[0]] + [repr(float(x)) for x in g[1:12]] + [g[12], repr(float(g[13]))]})
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(rows, f, ensure_ascii=True, indent=0)
    print(name, len(rows), "messages,", sum(r["vehicle"] is None for r in rows), "dropped")


def alerts_fixture(name):
    """AlertManager's table methods (warning_system.py:120-213, 259-285, 488-517) on the scripted scenario
    of tests/alert_cases.py, controlled clock: per op the events and the whole table."""
    import json
    from tests import alert_cases as A
    script = A.scenario()
    got = S.run_alert_scenario_A(script)
    rows = [{"events": [list(e) for e in ev], "table": [[t[0], t[1], repr(t[2]), repr(t[3]), t[4], repr(t[5]), t[6]] for t in tab]}
            for ev, tab in got]
    import gzip
    with gzip.GzipFile(os.path.join(HERE, name), "w", mtime=0) as f:
        f.write(json.dumps(rows).encode())
    print(name, len(rows), "ops,", sum(len(r["events"]) for r in rows), "events, final table", len(rows[-1]["table"]))


def helpers_fixture(name):
    """The per-pair helpers the reference's prediction model calls on the detector (collision_detection.py:296-389)
    and the alert message tiers (warning_system.py:313-329)."""
    import json
    frame = W.uniform_frame(500, 107, map_size=260.0, drone_fraction=0.3)
    rng = np.random.default_rng(5)
    f64 = W.frame_to_f64(frame)
    # pairs at every distance: neighbours within 30 m (most of them hit within the window) and random ones
    pos = np.stack([f64["px"], f64["py"], f64["pz"]], 1)
    pairs = []
    for i in range(0, 500, 3):
        d = np.linalg.norm(pos - pos[i], axis=1)
        near = [int(j) for j in np.argsort(d)[1:7]]
        pairs += [(i, j) for j in near] + [(i, int(j)) for j in rng.integers(0, 500, 2) if int(j) != i]
    out = {"frame": {k: v.tolist() for k, v in _frame_arrays(frame).items()}, "cases": []}
    for T in (10.0, 1.0, 0.35):
        res = S.run_pair_helpers_A(f64, pairs, T)
        out["cases"].append({"T": T, "pairs": pairs, "results": res})
        print(name, T, sum(r is not None for r in res), "hits of", len(res))
    msg_cases = [(float(r), f"vehicle-{k}", float(t), float(d)) for k, (r, t, d) in enumerate(
        (r, t, d) for r in (0.3, 0.45, 0.59999, 0.6, 0.79999, 0.8, 0.95, 1.0) for t in (0.0, 0.05, 0.14999, 2.25, 9.95, 19.4)
        for d in (0.0, 0.05, 0.25, 3.349999, 7.5, 12.25))]
    msg_cases.append((0.9, "车-7 \"x\"", 1.0, 2.0))
    out["messages"] = {"cases": msg_cases, "texts": S.run_alert_messages_A(msg_cases)}
    with open(os.path.join(HERE, name), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)


def partitioner_case(seed=5, n=300):
    """Scenario for the reference SpatialPartitioner (shared with tests/test_dropin_cpu.py)."""
    rng = np.random.default_rng(seed)
    f32 = lambda a: [float(np.float32(v)) for v in a]
    veh = list(zip(f32(rng.uniform(0, 6000, n)), f32(rng.uniform(0, 4000, n)), f32(rng.uniform(0, 100, n))))
    more = list(zip(f32(rng.uniform(5000, 9000, 40)), f32(rng.uniform(0, 4000, 40)), f32(rng.uniform(0, 100, 40))))
    queries = veh[:40] + more[:10] + [(7500.0, 100.0, 0.0), (-10.0, 5.0, 1.0), (999.5, 999.5, 99.5), (1000.0, 1000.0, 100.0)]
    ops = [["query"], ["stats"],
           ["loads", {"shard-0": 0.9, "shard-1": 0.1, "shard-2": 0.2, "shard-3": 0.95, "shard-4": 0.5, "shard-5": 0.5,
                      "shard-6": 0.25, "shard-7": 0.6}],
           ["rebalance"], ["query"], ["stats"],
           ["insert", more], ["query"],
           ["loads", {"shard-0": 0.2, "shard-1": 0.1, "shard-2": 0.2, "shard-3": 0.1, "shard-4": 0.75, "shard-5": 0.05,
                      "shard-6": 0.25, "shard-7": 0.99}],
           ["rebalance"], ["query"], ["stats"],
           ["loads", {"shard-4": 0.8, "shard-7": 0.71, "shard-0": 0.5, "shard-1": 0.5, "shard-2": 0.5, "shard-3": 0.5}],
           ["rebalance"], ["query"], ["stats"],
           ["loads", {f"shard-{k}": 0.1 for k in range(8)}],  # nothing overloaded, everything underloaded: merges
           ["rebalance"], ["query"], ["stats"], ["rebalance"], ["stats"]]
    return {"vehicles": veh, "num_shards": 8, "queries": queries, "ops": ops}


def partitioner_fixture(name):
    import json
    case = partitioner_case()
    res = S.run_partitioner_A(case["vehicles"], case["num_shards"], case["queries"], [tuple(o) for o in case["ops"]])
    case["results"] = res
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(case, f)
    print(name, [r for r in res if isinstance(r, dict) and "split_regions" in r])


def main():
    if not S.reference_available():
        raise SystemExit("needs /root/reference")
    # dense 3-D mixed vehicles + drones: many pairs inside the safe distance
    detect_fixture("detect_dense3d.npz", W.uniform_frame(700, 101, map_size=450.0, drone_fraction=0.3))
    # 2-D ground traffic, zero accelerations (quirk Q1 regime: survivors have cur <= safe)
    detect_fixture("detect_2d_noaccel.npz", W.uniform_frame(900, 102, map_size=600.0, accel=False))
    # non-default radius / window
    detect_fixture("detect_r60_t4.npz", W.uniform_frame(500, 103, map_size=300.0, drone_fraction=0.5),
                   R=60.0, T=4.0)
    # configs[0] generator (5 cities, 10 km map), first 1000 vehicles
    detect_fixture("detect_city1k.npz", W.reference_city_frame(1000, 1234))
    predict_fixture("predict_dense3d.npz", W.uniform_frame(350, 104, map_size=500.0, drone_fraction=0.3),
                    W.random_patterns(350, 7))
    predict_fixture("predict_2d_allcv.npz", W.uniform_frame(400, 105, map_size=700.0),
                    np.ones(400, np.uint8))
    predict_fixture("predict_city300.npz", W.take(W.reference_city_frame(3000, 1236), slice(0, 300)),
                    W.random_patterns(300, 8, p=(0.05, 0.45, 0.45, 0.05)))
    fb = W.uniform_frame(600, 106, map_size=160.0, drone_fraction=0.2)
    implB_fixture("implB_dense.npz", fb, np.random.default_rng(3).random(600) < 0.9)
    scalars_fixture("scalars.npz")
    ingest_fixture("ingest_messages.json")
    alerts_fixture("alert_scenario.json.gz")
    helpers_fixture("pair_helpers.json")
    partitioner_fixture("partitioner.json")


if __name__ == "__main__":
    main()
