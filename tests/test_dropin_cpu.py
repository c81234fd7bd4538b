"""Host-side logic of the drop-in classes that needs no GPU: API surface, id <-> slot staging,
truncating grid ids, alert classification."""
import inspect
import os

import numpy as np

from tests.helpers import GOLDEN


def test_api_surface_matches_reference_signatures():
    """Names and argument lists of SURVEY.md 8b (surface A and surface B)."""
    from rcd_b200.host import collision_detection as CD
    from rcd_b200.host import compute_node as CN
    from rcd_b200.host import spatial_index as SI
    from rcd_b200.host import warning_system as WS

    def params(fn):
        return [p for p in inspect.signature(fn).parameters if p not in ("self", "device")]

    assert params(SI.SpatialIndex.__init__) == ["base_size", "min_size", "max_level", "density_threshold_split",
                                                "density_threshold_merge", "adjustment_interval"]
    for name in ("get_cell_size", "get_grid_id", "get_grid_level", "insert_vehicle", "remove_vehicle",
                 "get_nearby_vehicles", "get_vehicle_position", "adjust_grid_resolution", "get_stats"):
        assert callable(getattr(SI.SpatialIndex, name))
    assert params(SI.SpatialPartitioner.__init__) == ["spatial_index", "num_shards", "min_load", "max_load",
                                                      "rebalance_interval"]
    for name in ("get_shard_for_position", "update_load", "check_rebalance", "rebalance_shards", "get_stats"):
        assert callable(getattr(SI.SpatialPartitioner, name))
    assert params(CD.CollisionDetector.__init__) == ["spatial_index"]
    assert params(CD.CollisionDetector.detect_collisions) == ["vehicle_id", "search_radius", "time_window"]
    sig = inspect.signature(CD.CollisionDetector.detect_collisions)
    assert sig.parameters["search_radius"].default == 100.0 and sig.parameters["time_window"].default == 10.0
    for name in ("update_vehicle", "remove_vehicle", "get_collision_risks", "get_stats", "_spatial_filtering",
                 "_predict_position"):
        assert callable(getattr(CD.CollisionDetector, name))
    assert params(CD.CollisionPredictionModel.__init__) == ["collision_detector"]
    assert params(CD.CollisionPredictionModel.update_trajectory) == ["vehicle_id", "position", "timestamp"]
    assert params(CD.CollisionPredictionModel.predict_collisions) == ["vehicle_id"]
    assert params(WS.AlertManager.process_collision_risks) == ["risks"]
    assert params(WS.AlertManager._get_priority) == ["risk_level", "time_to_collision"]
    assert params(CN.SpatialIndex.__init__) == ["cell_size"]
    for name in ("insert", "remove", "query_nearby", "get_position", "get_all_vehicles", "get_vehicle_count"):
        assert callable(getattr(CN.SpatialIndex, name))
    assert params(CN.VehicleState.__init__) == ["vehicle_id", "max_history"]
    assert params(CN.CollisionDetector.__init__) == ["prediction_time", "risk_threshold"]
    assert params(CN.CollisionDetector.detect_collisions) == ["vehicle", "nearby_vehicles"]


def test_object_table_keeps_dense_slots():
    from rcd_b200.host.object_table import ObjectTable
    t = ObjectTable(capacity=2)
    for k in range(7):
        t.set_state(f"v{k}", (k, 2 * k, 0), (1, 0, 0), (0, 0, 0), 0.5, 2.0, "car" if k % 2 else "bus")
    assert t.n == 7 and t.capacity >= 7 and t.type_codes == {"bus": 0, "car": 1}
    v0 = t.version
    assert t.remove("v2") and not t.remove("v2")
    assert t.n == 6 and t.version > v0
    assert sorted(t.slot_of) == ["v0", "v1", "v3", "v4", "v5", "v6"]
    assert sorted(t.slot_of.values()) == list(range(6))
    for vid, s in t.slot_of.items():
        assert t.ids[s] == vid and t.f["px"][s] == float(vid[1:]) and t.f["py"][s] == 2 * float(vid[1:])
    t.set_position("v9", 1.0, 2.0, 3.0)  # a bare position has zero velocity, like SpatialIndex.insert_vehicle
    s = t.slot_of["v9"]
    assert t.f["vx"][s] == 0 and t.f["pz"][s] == 3.0
    fr = t.frame()
    assert len(fr["px"]) == 7 and fr["type"].dtype == np.uint8


def test_grid_ids_truncate_toward_zero_like_the_reference():
    from rcd_b200.host.models import Position
    from rcd_b200.host.spatial_index import SpatialIndex
    z = np.load(os.path.join(GOLDEN, "scalars.npz"))
    idx = SpatialIndex.__new__(SpatialIndex)
    idx.base_size = (1000.0, 1000.0, 100.0)
    for lvl in range(4):
        got = np.array([idx.get_grid_id(Position(*p), lvl) for p in z["grid_pts"]])
        assert np.array_equal(got, z[f"grid_l{lvl}"])
    assert idx.get_grid_id(Position(-0.5, -999.9, -0.1), 0) == (0, 0, 0)


def test_alert_priorities_match_reference_golden():
    from rcd_b200.host.models import CollisionRisk, Position
    from rcd_b200.host.warning_system import AlertManager
    z = np.load(os.path.join(GOLDEN, "scalars.npz"))
    mgr = AlertManager()
    risks = []
    for k, ((risk, ttc), want) in enumerate(zip(z["prio_in"], z["prio_out"])):
        if want >= 0:
            assert mgr._get_priority(float(risk), float(ttc)) == want
        risks.append(CollisionRisk(id=str(k), vehicle_id="a", other_vehicle_id=f"b{k}", time_to_collision=float(ttc),
                                   distance=3.0, relative_speed=1.0, risk_level=float(risk),
                                   collision_position=Position(0, 0, 0), timestamp=0.0))
    alerts = mgr.process_collision_risks(risks)
    assert len(alerts) == int((z["prio_out"] >= 0).sum())  # risks below RISK_LEVEL_LOW raise no alert
    assert [a.priority for a in alerts] == [int(p) for p in z["prio_out"] if p >= 0]


def test_level0_grid_view_and_partitioner_on_host():
    from rcd_b200.host.models import Position
    from rcd_b200.host.spatial_index import SpatialIndex, SpatialPartitioner
    idx = SpatialIndex()
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 4000, (200, 2))
    for k, (x, y) in enumerate(pts):
        idx.insert_vehicle(f"v{k}", Position(float(x), float(y), 0.0))
    g = idx.grids
    assert sum(c.vehicle_count for c in g[0].values()) == 200 and g[1] == {}
    assert all(idx.get_grid_id(idx.get_vehicle_position(v), 0) == cell.grid_id for cell in g[0].values() for v in cell.vehicles)
    st = idx.get_stats()
    assert st["total_vehicles"] == 200 and st["levels"][0]["vehicles"] == 200
    part = SpatialPartitioner(idx, num_shards=4)
    shards = [part.get_shard_for_position(Position(float(x), float(y), 0.0)) for x, y in pts]
    assert set(shards) <= {"shard-0", "shard-1", "shard-2", "shard-3"}
    assert shards == [f"shard-{hash(idx.get_grid_id(Position(float(x), float(y), 0.0), 0)) % 4}" for x, y in pts]
    assert part.get_shard_for_position(Position(9500.0, 100.0, 0.0)) is None  # no region there (spatial_index.py:521)
    idx.remove_vehicle("v0")
    assert idx.get_vehicle_position("v0") is None and sum(c.vehicle_count for c in idx.grids[0].values()) == 199


def _canon_shard(part, shard, num_shards):
    if shard is None:
        return None
    k = list(part.shard_loads).index(shard)
    return shard if k < num_shards else f"new-{k - num_shards}"


def test_partitioner_region_model_matches_reference_golden():
    """SpatialPartitioner in region mode against the reference's own class (tests/golden/partitioner.json, made by
    oracle/ref_shim.run_partitioner_A): initial hash placement, split of overloaded shards' regions (cells that were
    split then answer None: the level-0 lookup never descends), merge of adjacent single-cell regions, statistics."""
    import json
    from rcd_b200.host.models import Position
    from rcd_b200.host.spatial_index import SpatialIndex, SpatialPartitioner
    case = json.load(open(os.path.join(GOLDEN, "partitioner.json")))
    idx = SpatialIndex(adjustment_interval=float("inf"))
    n_ins = 0
    for x, y, z in case["vehicles"]:
        idx.insert_vehicle(f"v{n_ins}", Position(x, y, z))
        n_ins += 1
    ns = case["num_shards"]
    part = SpatialPartitioner(idx, num_shards=ns)
    n_split = n_merge = 0
    for op, want in zip(case["ops"], case["results"]):
        if op[0] == "query":
            got = [_canon_shard(part, part.get_shard_for_position(Position(x, y, z)), ns) for x, y, z in case["queries"]]
            assert got == want
        elif op[0] == "loads":
            names = list(part.shard_loads)
            for name, load in op[1].items():
                part.update_load(names[ns + int(name[4:])] if name.startswith("new-") else name, load)
        elif op[0] == "rebalance":
            r = part.rebalance_shards()
            assert {k: v for k, v in r.items() if k != "elapsed_ms"} == want
            n_split += r["split_regions"]
            n_merge += r["merged_regions"]
        elif op[0] == "insert":
            for x, y, z in op[1]:
                idx.insert_vehicle(f"v{n_ins}", Position(x, y, z))
                n_ins += 1
        elif op[0] == "stats":
            st = part.get_stats()
            assert st["total_shards"] == want["total_shards"] and st["total_regions"] == want["total_regions"]
            assert {_canon_shard(part, s, ns): v for s, v in st["shards"].items()} == want["shards"]
            regions = sorted((sorted([lvl, list(g)] for lvl, g in cells), _canon_shard(part, part.region_to_shard.get(rid), ns))
                             for rid, cells in part.regions.items())
            assert [[r[0], r[1]] for r in regions] == [[w[0], w[1]] for w in want["regions"]]
    assert n_split >= 4 and n_merge >= 1


def test_shard_manager_sticky_and_slab_routing():
    """get_shard_for_vehicle (data_sharding.py:172-201): sticky + random fallback in region mode; in slab mode the
    owner follows the position (migration) and re-balancing moves the cuts towards the slower slab."""
    import random
    from rcd_b200.host.data_sharding import ShardManager
    from rcd_b200.host.models import Position
    from rcd_b200.host.spatial_index import SpatialIndex, SpatialPartitioner
    from rcd_b200.host import slabs as S
    idx = SpatialIndex()
    part = SpatialPartitioner(idx, num_shards=4)  # built on an empty index like collision_system.py:171-180: no regions
    mgr = ShardManager(part, initial_shards=4, rng=random.Random(3))
    first = mgr.get_shard_for_vehicle("a", Position(10.0, 10.0, 0.0))
    assert first in mgr.shards  # random fallback (:190-192)
    assert all(mgr.get_shard_for_vehicle("a", Position(float(x), 0.0, 0.0)) == first for x in (0, 5000, 9000))  # sticky
    assert mgr.shards[first]["vehicle_count"] == 1 and mgr.get_stats()["total_vehicles"] == 1
    # slab mode
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0, 2000, 3000), rng.uniform(2000, 8000, 1000)]).astype(np.float32)
    for k, x in enumerate(xs):
        idx.insert_vehicle(f"v{k}", Position(float(x), float(rng.uniform(0, 8000)), 0.0))
    frame = {"px": xs, "py": np.zeros_like(xs)}
    lo, hi = S.slab_bounds(frame, 4, 8000.0, pair_weight=0.0)

    class _Ex:
        cuts = None

        def set_cuts(self, lo, hi):
            self.cuts = (lo, hi)

    ex = _Ex()
    part.attach_slabs(lo, hi, 8000.0, exchanges=[ex])
    mgr2 = ShardManager(part, initial_shards=4)
    own = [mgr2.get_shard_for_vehicle(f"v{k}", Position(float(x), 0.0, 0.0)) for k, x in enumerate(xs)]
    assert own == [f"shard-{o}" for o in S.owner_of(xs, lo, hi)]
    assert sum(v["vehicle_count"] for v in mgr2.shards.values()) == len(xs) and mgr2.migrations == 0
    # a vehicle that crosses a cut migrates
    x_new = float(hi[0]) + 1.0
    k0 = int(np.argmin(xs))
    before = dict((s, v["vehicle_count"]) for s, v in mgr2.shards.items())
    assert mgr2.get_shard_for_vehicle(f"v{k0}", Position(x_new, 0.0, 0.0)) == "shard-1" and mgr2.migrations == 1
    assert mgr2.shards["shard-0"]["vehicle_count"] == before["shard-0"] - 1
    assert mgr2.shards["shard-1"]["vehicle_count"] == before["shard-1"] + 1
    # slab 0 reports the longest frame time: its cut moves left, the exchange follows
    mgr2.update_shard_loads([9.0, 3.0, 3.0, 3.0])
    r = part.rebalance_shards()
    assert r["cuts_moved"] and ex.cuts is not None and float(part.slab_cuts[1][0]) < float(hi[0])
    st = part.get_stats()
    assert sum(v["vehicles"] for v in st["shards"].values()) == len(xs) and len(st["slab_bounds"]) == 4
