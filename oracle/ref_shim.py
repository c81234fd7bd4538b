"""TEST INFRASTRUCTURE ONLY -- runs the reference's own ``src/collision`` bytecode under a
minimal in-memory shim, to pin the oracle and to generate ``tests/golden`` fixtures.

The reference (``/root/reference``) is pure Python and its implementation A cannot be
imported as shipped.  This module applies the documented repairs
R1-R5 of SURVEY.md section 8(c) *in memory* (no reference file is touched or copied) and then
drives the reference's own classes:

  R1  add the missing ``Vehicle`` / ``Alert`` dataclasses to ``src.common.models``
      (fields per src/collision/warning_system.py:649-670).
  R2  rebind ``models.CollisionRisk`` to the 10-field record the detector really constructs
      (src/collision/collision_detection.py:156-166, 831-842); done after importing
      ``src.compute.compute_node`` which needs the original 9-field one.
  R3  ``Timer.elapsed_ms`` usable both as a number and as a call (src/common/utils.py:56-58 vs
      collision_detection.py:180-185).
  R4  keep ``vehicle_positions[id]`` after ``SpatialIndex.insert_vehicle`` (the shipped code
      wipes it, src/collision/spatial_index.py:171,178,225-227).
  R5  ``SpatialIndex(adjustment_interval=inf)`` -> level-0 grid only (section 8a, a2).

It only exists in the build container: ``/root/reference`` is absent on the GPU box, so
nothing at run time (tests -m gpu, smoke, bench) may import this file.  It is used by
``tests/golden/make_golden.py`` and by the ``needs_reference`` CPU tests.
"""
from __future__ import annotations

import dataclasses
import math
import os
import sys
import types
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

REFERENCE_ROOT = os.environ.get("RCD_REFERENCE_ROOT", "/root/reference")

_loaded: Optional[types.SimpleNamespace] = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "collision", "collision_detection.py"))


class _ElapsedMs(float):
    """R3: a float that can also be called (``timer.elapsed_ms`` and ``timer.elapsed_ms()``)."""

    def __call__(self) -> float:  # pragma: no cover - trivial
        return float(self)


def load_reference() -> types.SimpleNamespace:
    """Import the reference modules under the shim; returns a namespace of its classes."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import logging

    logging.disable(logging.CRITICAL)  # the reference logs at INFO on every constructor

    import src.common.models as models  # type: ignore
    import src.common.utils as utils  # type: ignore

    # impl. B needs the ORIGINAL CollisionRisk: import it before R2.
    import src.compute.compute_node as compute_node  # type: ignore

    Position, Vector = models.Position, models.Vector

    # --- R1 -----------------------------------------------------------------------------
    @dataclass
    class Vehicle:
        id: str
        position: Any
        velocity: Any
        acceleration: Any
        heading: float
        size: float
        type: str
        timestamp: float

    @dataclass
    class Alert:
        id: str = ""
        vehicle_id: str = ""
        other_vehicle_id: str = ""
        risk_level: float = 0.0
        time_to_collision: float = 0.0
        message: str = ""
        priority: int = 0
        timestamp: float = 0.0

    models.Vehicle = Vehicle
    models.Alert = Alert

    # --- R2 -----------------------------------------------------------------------------
    @dataclass
    class CollisionRiskA:
        id: str
        vehicle_id: str
        other_vehicle_id: str
        time_to_collision: float
        distance: float
        relative_speed: float
        risk_level: float
        collision_position: Any
        timestamp: float
        is_predicted: bool = False

    original_risk = models.CollisionRisk
    models.CollisionRisk = CollisionRiskA

    # --- R3 -----------------------------------------------------------------------------
    def _elapsed_ms(self):
        return _ElapsedMs(self.elapsed() * 1000.0)

    utils.Timer.elapsed_ms = property(_elapsed_ms)

    import src.collision.spatial_index as spatial_index  # type: ignore
    import src.collision.collision_detection as collision_detection  # type: ignore
    import src.collision.warning_system as warning_system  # type: ignore

    # --- R4 -----------------------------------------------------------------------------
    _orig_insert = spatial_index.SpatialIndex.insert_vehicle

    def _insert_vehicle(self, vehicle_id, position):
        _orig_insert(self, vehicle_id, position)
        self.vehicle_positions[vehicle_id] = position
        self.stats["total_vehicles"] = len(self.vehicle_positions)

    spatial_index.SpatialIndex.insert_vehicle = _insert_vehicle

    models.CollisionRisk = CollisionRiskA  # what src.collision.* bound at import
    _loaded = types.SimpleNamespace(
        models=models,
        utils=utils,
        Position=Position,
        Vector=Vector,
        Vehicle=Vehicle,
        CollisionRiskA=CollisionRiskA,
        CollisionRiskB=original_risk,
        spatial_index=spatial_index,
        collision_detection=collision_detection,
        warning_system=warning_system,
        compute_node=compute_node,
    )
    return _loaded


# ----------------------------------------------------------------------------------------
# Frame drivers.  A frame is a dict of 1-D arrays (see oracle/frames.py):
#   px py pz vx vy vz ax ay az size heading : float (fp32-representable values)
#   type : small int (mapped to a distinct string per value)
# ----------------------------------------------------------------------------------------

def _vehicles_from_frame(ref, frame) -> List[Any]:
    n = len(frame["px"])
    out = []
    for i in range(n):
        out.append(
            ref.Vehicle(
                id=f"v{i}",
                position=ref.Position(float(frame["px"][i]), float(frame["py"][i]), float(frame["pz"][i])),
                velocity=ref.Vector(float(frame["vx"][i]), float(frame["vy"][i]), float(frame["vz"][i])),
                acceleration=ref.Vector(float(frame["ax"][i]), float(frame["ay"][i]), float(frame["az"][i])),
                heading=float(frame["heading"][i]),
                size=float(frame["size"][i]),
                type=f"type{int(frame['type'][i])}",
                timestamp=0.0,
            )
        )
    return out


def build_detector_A(frame):
    """SpatialIndex(level-0 only, R5) + CollisionDetector loaded with the frame."""
    ref = load_reference()
    index = ref.spatial_index.SpatialIndex(adjustment_interval=math.inf)
    det = ref.collision_detection.CollisionDetector(index)
    vehicles = _vehicles_from_frame(ref, frame)
    for v in vehicles:
        det.update_vehicle(v)
    return ref, index, det, vehicles


def run_detect_A(frame, search_radius: float = 100.0, time_window: float = 10.0) -> Dict[str, Any]:
    """Run reference ``CollisionDetector.detect_collisions`` for every vehicle.

    Returns per-stage tables in dense-index space:
      candidates : sorted list of directed (i, j)            (stage 1, collision_detection.py:208-227)
      potentials : sorted list of (i, j, tc, cd)             (stage 2, :229-294)
      risks      : sorted list of (i, j, ttc, distance, rel_speed, risk, cx, cy, cz)
    """
    ref, index, det, vehicles = build_detector_A(frame)
    n = len(vehicles)
    cands: List[Tuple[int, int]] = []
    pots: List[Tuple[int, int, float, float]] = []
    risks: List[Tuple] = []
    for i in range(n):
        vid = f"v{i}"
        near = det._spatial_filtering(vid, vehicles[i].position, search_radius)
        nearby = {o: det.vehicle_cache[o] for o in near}
        for o in near:
            cands.append((i, int(o[1:])))
        for (o, tc, cd) in det._temporal_filtering(vehicles[i], nearby, time_window):
            pots.append((i, int(o[1:]), tc, cd))
        for r in det.detect_collisions(vid, search_radius, time_window):
            p = r.collision_position
            risks.append((i, int(r.other_vehicle_id[1:]), r.time_to_collision, r.distance,
                          r.relative_speed, r.risk_level, p.x, p.y, p.z))
    cands.sort()
    pots.sort()
    risks.sort()
    return {"candidates": cands, "potentials": pots, "risks": risks, "stats": dict(det.stats)}


PATTERN_NAMES = ("stationary", "constant_velocity", "accelerating")


def run_predict_A(frame, pattern: Sequence[int]) -> Dict[str, Any]:
    """Run reference ``CollisionPredictionModel.predict_collisions`` for every vehicle.

    ``pattern[i]`` in {0 stationary, 1 constant_velocity, 2 accelerating, 3 no-history}.  The
    pattern classifier is exercised separately (``run_pattern_A``); here
    ``_analyze_trajectory_pattern`` is overridden per vehicle so the offsets/merge logic
    (collision_detection.py:713-865) runs the reference's own bytecode for a chosen class.
    3 leaves the history empty -> the reference falls back to detect_collisions (:590-592).
    """
    ref, index, det, vehicles = build_detector_A(frame)
    model = ref.collision_detection.CollisionPredictionModel(det)
    n = len(vehicles)
    current = {"i": 0}

    def _pattern(_history):
        p = int(pattern[current["i"]])
        return {"type": PATTERN_NAMES[p] if p < 3 else "unknown"}

    model._analyze_trajectory_pattern = _pattern
    for i in range(n):
        if int(pattern[i]) != 3:
            model.update_trajectory(f"v{i}", vehicles[i].position, 0.0)
            model.update_trajectory(f"v{i}", vehicles[i].position, 1.0)
    risks: List[Tuple] = []
    for i in range(n):
        current["i"] = i
        for r in model.predict_collisions(f"v{i}"):
            p = r.collision_position
            risks.append((i, int(r.other_vehicle_id[1:]), float(r.time_to_collision), r.distance,
                          r.relative_speed, r.risk_level, p.x, p.y, p.z,
                          bool(getattr(r, "is_predicted", False))))
    risks.sort()
    return {"risks": risks}


def run_pattern_A(histories: Sequence[Sequence[Tuple[float, float, float, float]]]) -> List[int]:
    """Reference ``_analyze_trajectory_pattern`` (collision_detection.py:623-711) on explicit
    histories [(x, y, z, t), ...] -> class index (0/1/2), or 3 for 'unknown' (< 2 samples)."""
    ref = load_reference()
    model = ref.collision_detection.CollisionPredictionModel.__new__(
        ref.collision_detection.CollisionPredictionModel)
    out = []
    for h in histories:
        hist = [(ref.Position(x, y, z), t) for (x, y, z, t) in h]
        name = model._analyze_trajectory_pattern(hist)["type"]
        out.append(PATTERN_NAMES.index(name) if name in PATTERN_NAMES else 3)
    return out


def run_priority_A(risk_ttc: Sequence[Tuple[float, float]]) -> List[int]:
    """Reference alert gate + ``AlertManager._get_priority`` (warning_system.py:259-311).
    Returns -1 where the risk is below RISK_LEVEL_LOW (no alert)."""
    ref = load_reference()
    ws = ref.warning_system
    mgr = ws.AlertManager.__new__(ws.AlertManager)
    out = []
    for risk, ttc in risk_ttc:
        if risk < ws.RISK_LEVEL_LOW:
            out.append(-1)
        else:
            out.append(int(mgr._get_priority(risk, ttc)))
    return out


def run_alert_messages_A(cases: Sequence[Tuple[float, str, float, float]]) -> List[str]:
    """``AlertManager._generate_alert_message`` (warning_system.py:313-329) for (risk_level, other_vehicle_id,
    time_to_collision, distance) tuples."""
    ref = load_reference()
    ws = ref.warning_system
    mgr = ws.AlertManager.__new__(ws.AlertManager)
    out = []
    for risk, other, ttc, dist in cases:
        r = types.SimpleNamespace(risk_level=risk, other_vehicle_id=other, time_to_collision=ttc, distance=dist)
        out.append(mgr._generate_alert_message(r))
    return out


def run_pair_helpers_A(frame, pairs: Sequence[Tuple[int, int]], time_window: float) -> List[Any]:
    """``CollisionDetector._precise_collision_detection`` (:296-342) and ``_risk_assessment`` (:344-389) for the
    listed pairs of a frame: per pair None, or (collision_time, distance, safe_distance, relative_speed,
    mid x, y, z, risk)."""
    ref, _index, det, vehicles = build_detector_A(frame)
    out = []
    for i, j in pairs:
        info = det._precise_collision_detection(vehicles[i], vehicles[j], time_window)
        if info is None:
            out.append(None)
            continue
        risk = det._risk_assessment(vehicles[i], vehicles[j], info)
        m = info["collision_position"]
        out.append((info["collision_time"], info["distance"], info["safe_distance"], info["relative_speed"], m.x, m.y, m.z, risk))
    return out


def run_nearby_A(frame, queries: Sequence[Tuple[float, float, float]], radius: float) -> List[List[int]]:
    """Reference ``SpatialIndex.get_nearby_vehicles`` (spatial_index.py:229-271) for explicit
    query points (self is NOT stripped: Q8)."""
    ref, index, det, vehicles = build_detector_A(frame)
    out = []
    for (x, y, z) in queries:
        ids = index.get_nearby_vehicles(ref.Position(x, y, z), radius)
        out.append(sorted(int(s[1:]) for s in ids))
    return out


def run_grid_id_A(points: Sequence[Tuple[float, float, float]], level: int = 0) -> List[Tuple[int, int, int]]:
    """Reference ``SpatialIndex.get_grid_id`` (spatial_index.py:97-112): trunc-toward-zero ids."""
    ref = load_reference()
    index = ref.spatial_index.SpatialIndex(adjustment_interval=math.inf)
    return [tuple(index.get_grid_id(ref.Position(x, y, z), level)) for (x, y, z) in points]


def run_B(frame, has_history: Sequence[bool], radius: float = 100.0) -> Dict[str, Any]:
    """Reference impl. B (src/compute/compute_node.py:20-321) unmodified: index every vehicle,
    then for each vehicle query_nearby(radius) + CollisionDetector.detect_collisions.

    ``has_history[i]`` False -> the vehicle has a single LocationData sample, so
    ``predict_position`` returns None and every pair involving it is skipped (:202-203,269-270).
    Returns (i, j, risk, ttc_formula, rel_speed, cx, cy, cz); ttc_formula is recomputed from the
    pair's distances because the dataclass field is wall-clock perturbed (quirk Q10)."""
    ref = load_reference()
    cn = ref.compute_node
    n = len(frame["px"])
    index = cn.SpatialIndex()
    states = {}
    for i in range(n):
        vid = f"v{i}"
        st = cn.VehicleState(vid)
        loc = ref.models.LocationData(
            vehicle_id=vid, timestamp=0.0,
            position=ref.Position(float(frame["px"][i]), float(frame["py"][i]), float(frame["pz"][i])),
            velocity=ref.Vector(float(frame["vx"][i]), float(frame["vy"][i]), float(frame["vz"][i])),
            heading=float(frame["heading"][i]), vehicle_type=f"type{int(frame['type'][i])}")
        st.update(loc)
        if has_history[i]:
            st.update(loc)
        states[vid] = st
        index.insert(vid, loc.position)
    det = cn.CollisionDetector()
    cands: List[Tuple[int, int]] = []
    risks: List[Tuple] = []
    import time as _time
    for i in range(n):
        vid = f"v{i}"
        st = states[vid]
        loc = st.get_current_location()
        near = index.query_nearby(loc.position, radius)
        for o in near:
            cands.append((i, int(o[1:])))
        nearby = {o: states[o] for o in near}
        t0 = _time.time()
        for r in det.detect_collisions(st, nearby):
            # estimated_collision_time = time.time() + ttc  (compute_node.py:314) -> recover ttc
            # to ~1e-6 s; the exact formula value is re-derived by the oracle comparison.
            ttc_est = r.estimated_collision_time - r.timestamp
            risks.append((i, int(r.vehicle_id2[1:]), r.risk_level, ttc_est, r.relative_velocity,
                          r.position.x, r.position.y, r.position.z))
        del t0
    cands.sort()
    risks.sort()
    return {"candidates": cands, "risks": risks}


# ----------------------------------------------------------------------------------------
# Ingest driver: EarlyWarningSystem._handle_vehicle_position (warning_system.py:638-678) on raw
# message texts.  One more repair is needed to execute it at all:
#   R7  warning_system uses `Vector` without importing it (:656) -> bind models.Vector there.
# The handler is run on an instance created without __init__ (its constructor needs the broker),
# with recording stand-ins for the two objects it calls.
# ----------------------------------------------------------------------------------------
def run_handle_position_A(texts: Sequence[str]) -> List[Any]:
    """For every message text: None if the reference drops it, else the tuple
    (id, px, py, pz, vx, vy, vz, ax, ay, az, heading, size, type, timestamp) of the Vehicle it
    hands to update_vehicle, after checking that update_trajectory got the same position/time."""
    import json

    ref = load_reference()
    ws = ref.warning_system
    ws.Vector = ref.Vector  # R7

    class _Rec:
        def __init__(self):
            self.vehicle = None
            self.traj = None

        def update_vehicle(self, v):
            self.vehicle = v

        def update_trajectory(self, vid, pos, ts):
            self.traj = (vid, pos, ts)

    class _Msg:
        def __init__(self, value):
            self.value = value

    out: List[Any] = []
    for t in texts:
        rec = _Rec()
        sys_ = object.__new__(ws.EarlyWarningSystem)
        sys_.collision_detector = rec
        sys_.prediction_model = rec
        try:
            value = json.loads(t)  # what the broker hands over (messaging.py deserialises with json)
        except Exception:
            out.append(None)
            continue
        sys_._handle_vehicle_position(_Msg(value))
        v = rec.vehicle
        if v is None or rec.traj is None:
            out.append(None)
            continue
        assert rec.traj[0] == v.id and rec.traj[1] is v.position and rec.traj[2] is v.timestamp
        out.append((v.id, v.position.x, v.position.y, v.position.z, v.velocity.x, v.velocity.y, v.velocity.z,
                    v.acceleration.x, v.acceleration.y, v.acceleration.z, v.heading, v.size, v.type, v.timestamp))
    return out


# ----------------------------------------------------------------------------------------
# Alert lifecycle driver: the reference's AlertManager table logic (warning_system.py:120-213,
# 259-285, 488-517) on a scripted scenario with a controlled clock.  The manager is created
# without __init__ (which needs the broker); only the three dict/list attributes its table methods
# use are set, exactly as __init__ sets them (:61-66).
# ----------------------------------------------------------------------------------------
def run_alert_scenario_A(script) -> List[Any]:
    """script: list of ("process", now, [(vid, other, risk, ttc, distance)]) | ("ack", [(vid, other)]) |
    ("cleanup", now).  Returns, per op, (events, table) with events = sorted [(kind, vid, other, priority,
    old_priority)] and table = sorted [(vid, other, risk, ttc, priority, timestamp, acknowledged)]."""
    ref = load_reference()
    ws = ref.warning_system

    class _Clock:
        now = 0.0

        @staticmethod
        def time():
            return _Clock.now

    real_time = ws.time
    ws.time = _Clock
    try:
        am = object.__new__(ws.AlertManager)
        am.alerts, am.vehicle_alerts, am.alert_queue = {}, {}, []
        out = []
        for op in script:
            events = []
            if op[0] == "process":
                _Clock.now = op[1]
                before = {(a.vehicle_id, a.other_vehicle_id): (a.id, a.priority) for a in am.alerts.values()}
                risks = [ref.CollisionRiskA(id=f"r{k}", vehicle_id=v, other_vehicle_id=o, time_to_collision=ttc, distance=d,
                                            relative_speed=0.0, risk_level=r, collision_position=None, timestamp=op[1])
                         for k, (v, o, r, ttc, d) in enumerate(op[2])]
                got = am.process_collision_risks(risks)
                for a in got:
                    key = (a.vehicle_id, a.other_vehicle_id)
                    if key not in before or before[key][0] != a.id:
                        events.append(("created", key[0], key[1], a.priority, -1))
                    else:
                        old = before[key][1]
                        events.append(("changed" if a.priority != old else "refreshed", key[0], key[1], a.priority, old))
                # a priority change must have re-queued the alert (:186-191); every alert is queued once
                assert sorted(x.id for x in am.alert_queue) == sorted(am.alerts)
            elif op[0] == "ack":
                for v, o in op[1]:
                    aid = am.vehicle_alerts.get(v, {}).get(o)
                    if aid is not None:
                        am.acknowledge_alert(aid)
            elif op[0] == "cleanup":
                _Clock.now = op[1]
                before = {(a.vehicle_id, a.other_vehicle_id) for a in am.alerts.values()}
                am._cleanup_expired_alerts()
                after = {(a.vehicle_id, a.other_vehicle_id) for a in am.alerts.values()}
                events = [("expired", v, o, -1, -1) for v, o in before - after]
            table = sorted((a.vehicle_id, a.other_vehicle_id, a.risk_level, a.time_to_collision, a.priority, a.timestamp,
                            bool(a.acknowledged)) for a in am.alerts.values())
            out.append((sorted(events), table))
        return out
    finally:
        ws.time = real_time


# ----------------------------------------------------------------------------------------
# Partitioner driver: the reference's SpatialPartitioner (spatial_index.py:435-862) on a scripted scenario.
# R6: __init__ calls _initialize_shards() before it creates self.stats (:466-469 vs :493); the shim gives the
# instance an empty stats dict first.  Region ids and new shard ids are uuids: they are reported by canonical
# names (a region by its sorted cells, a new shard as "new-<k>" in order of creation).
# ----------------------------------------------------------------------------------------
def run_partitioner_A(vehicles: Sequence[Tuple[float, float, float]], num_shards: int, queries, ops) -> List[Any]:
    """ops: ("query",) | ("loads", {shard_name: load}) | ("rebalance",) | ("stats",) | ("insert", [(x, y, z)]).
    Returns one result per op."""
    ref = load_reference()
    si = ref.spatial_index
    index = si.SpatialIndex(adjustment_interval=math.inf)
    n_ins = 0
    for (x, y, z) in vehicles:
        index.insert_vehicle(f"v{n_ins}", ref.Position(x, y, z))
        n_ins += 1
    orig_init = si.SpatialPartitioner._initialize_shards

    def _init_with_stats(self):
        self.stats = {}
        orig_init(self)

    si.SpatialPartitioner._initialize_shards = _init_with_stats
    try:
        part = si.SpatialPartitioner(index, num_shards=num_shards)
    finally:
        si.SpatialPartitioner._initialize_shards = orig_init

    def canon(shard):
        if shard is None:
            return None
        names = list(part.shard_loads)
        k = names.index(shard)
        return shard if k < num_shards else f"new-{k - num_shards}"

    out: List[Any] = []
    for op in ops:
        if op[0] == "query":
            out.append([canon(part.get_shard_for_position(ref.Position(x, y, z))) for (x, y, z) in queries])
        elif op[0] == "loads":
            names = list(part.shard_loads)
            for name, load in op[1].items():
                real = names[num_shards + int(name[4:])] if name.startswith("new-") else name
                part.update_load(real, load)
            out.append(None)
        elif op[0] == "rebalance":
            r = part.rebalance_shards()
            out.append({k: v for k, v in r.items() if k != "elapsed_ms"})
        elif op[0] == "insert":
            for (x, y, z) in op[1]:
                index.insert_vehicle(f"v{n_ins}", ref.Position(x, y, z))
                n_ins += 1
            out.append(None)
        elif op[0] == "stats":
            st = part.get_stats()
            regions = sorted((sorted((lvl, tuple(g)) for lvl, g in cells), canon(part.region_to_shard.get(rid)))
                             for rid, cells in part.regions.items())
            out.append({"total_shards": st["total_shards"], "total_regions": st["total_regions"],
                        "shards": {canon(s): {"regions": v["regions"], "vehicles": v["vehicles"], "load": v["load"]}
                                   for s, v in st["shards"].items()},
                        "regions": regions})
    return out
