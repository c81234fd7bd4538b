"""Message texts for the ingest tests: the reference producer's format (vehicle_simulator.py:721-752)
on seeded random vehicles, plus hand-written edge cases.  Shared by tests/golden/make_golden.py
(which runs the reference's handler on them) and tests/test_ingest_cpu.py."""
import json
import random


def reference_message(rng: random.Random, k: int, timestamp: float, vid=None) -> str:
    """Exactly what VehicleSimulator.get_vehicle_json emits (json.dumps of the nested dict)."""
    vtype = rng.choice(["car", "truck", "bus", "motorcycle", "drone"])
    return json.dumps({
        "id": vid if vid is not None else f"vehicle-{k}",
        "position": {"x": rng.uniform(0, 10000), "y": rng.uniform(0, 10000), "z": rng.choice([0.0, rng.uniform(0, 100)])},
        "velocity": {"x": rng.uniform(-30, 30), "y": rng.uniform(-30, 30), "z": rng.choice([0.0, rng.uniform(-5, 5)])},
        "acceleration": {"x": rng.uniform(-1, 1), "y": rng.uniform(-1, 1), "z": 0.0},
        "heading": rng.uniform(0, 6.28318),
        "size": {"car": 2.0, "truck": 4.0, "bus": 5.0, "motorcycle": 1.0, "drone": rng.uniform(1, 5)}[vtype],
        "type": vtype,
        "timestamp": timestamp,
    })


def seeded_messages(n: int, seed: int):
    rng = random.Random(seed)
    return [reference_message(rng, k, 1_700_000_000.0 + 0.001 * k) for k in range(n)]


_BASE = {"id": "v", "position": {"x": 1.0, "y": 2.0, "z": 3.0}, "velocity": {"x": 4.0, "y": 5.0, "z": 6.0},
         "acceleration": {"x": 7.0, "y": 8.0, "z": 9.0}, "heading": 0.5, "size": 2.0, "type": "car", "timestamp": 10.0}


def _with(**kw):
    d = json.loads(json.dumps(_BASE))
    d.update(kw)
    return d


def edge_messages():
    """(text, note) pairs; every text is a single line."""
    cases = []
    add = lambda obj, note: cases.append((obj if isinstance(obj, str) else json.dumps(obj), note))
    add(_with(), "plain")
    add(_with(id="车辆-7 \"quoted\" \\ / \t tab"), "escapes and non-ASCII id (ensure_ascii -> \\uXXXX)")
    add(json.dumps(_with(id="émoji-\U0001F697"), ensure_ascii=False), "raw UTF-8 id")
    add(_with(id="surrogate-\U0001F697"), "surrogate pair escape")
    add(_with(position={"z": -0.0, "y": 1e-320, "x": 1.7976931348623157e308}), "key order, denormal, max double, -0.0")
    add(_with(velocity={"x": 1, "y": -2, "z": 0}), "integers")
    add(_with(heading=1E+2, size=2.5e-1), "exponent forms")
    add(_with(timestamp=float("nan")), "NaN timestamp (json.dumps emits NaN)")
    add(_with(position={"x": float("inf"), "y": float("-inf"), "z": 0.0}), "Infinity")
    add(_with(extra={"nested": [1, 2, {"a": None}], "s": "x}"}, other=True), "unknown keys are ignored")
    add('{"id": "dup", "id": "dup2", "position": {"x": 1, "x": 5, "y": 2, "z": 3}, "velocity": {"x": 4, "y": 5, "z": 6}, '
        '"acceleration": {"x": 7, "y": 8, "z": 9}, "heading": 0.5, "size": 2, "type": "car", "timestamp": 10, "size": 3}',
        "duplicate keys: last wins")
    add('  {  "id" : "ws" ,"position":{"x":1,"y":2,"z":3},"velocity":{"x":4,"y":5,"z":6},\t"acceleration":{"x":7,"y":8,"z":9},'
        '"heading":0.5,"size":2,"type":"car","timestamp":10}  ', "whitespace")
    add(_with(type=""), "empty type string")
    add(_with(id=""), "empty id")
    # dropped by the reference (KeyError / TypeError inside the handler)
    d = _with(); del d["heading"]; add(d, "missing heading")
    d = _with(); del d["position"]["z"]; add(d, "missing position.z")
    d = _with(); del d["id"]; add(d, "missing id")
    add(_with(position=[1, 2, 3]), "position is a list")
    add(_with(velocity=None), "velocity is null")
    add('{"id": "cut", "position": {"x": 1', "truncated")
    add('{"id": "bad", "position": {"x": 01, "y": 2, "z": 3}}', "leading zero is not JSON")
    add('not json at all', "garbage")
    add('{"id": "ctl\x01", "position": {"x": 1, "y": 2, "z": 3}}', "raw control character in a string")
    add('[1, 2, 3]', "not an object")
    return cases
