"""Batched ingest on the GPU (csrc/rcd_ingest.cuh): message buffers -> frame state + trajectory rings
must leave the device in exactly the state the per-field upload / history calls produce."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _messages(frame, ids, timestamp, order=None):
    """The reference producer's JSON (vehicle_simulator.py:721-752) for the objects of a frame."""
    types = ["car", "truck", "bus", "motorcycle", "drone", "tram", "van", "bike"]
    out = []
    for i in (range(len(ids)) if order is None else order):
        f = lambda k: float(frame[k][i])
        out.append(json.dumps({
            "id": ids[i], "position": {"x": f("px"), "y": f("py"), "z": f("pz")},
            "velocity": {"x": f("vx"), "y": f("vy"), "z": f("vz")},
            "acceleration": {"x": f("ax"), "y": f("ay"), "z": f("az")},
            "heading": f("heading"), "size": f("size"), "type": types[int(frame["type"][i]) % len(types)],
            "timestamp": timestamp}))
    return "\n".join(out)


def test_messages_give_the_same_frame_as_upload():
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    from rcd_b200.host.ingest import VehiclePositionStream
    frame = W.uniform_frame(3000, 31, map_size=900.0, drone_fraction=0.3)
    ids = [f"vehicle-{k}" for k in range(3000)]
    rng = np.random.default_rng(1)
    order = rng.permutation(3000)  # messages arrive in any order: slot = order of first appearance
    with VehiclePositionStream(4096, 1 << 20) as s, FrameEngine(4096, 1 << 20) as e:
        assert s.handle_messages(_messages(frame, ids, 12.5, order)) == 3000
        got = s.detect_all_vehicles(predict=False)
        slot = np.array([s.ingest.slot_of(v) for v in ids])
        assert np.array_equal(slot[order], np.arange(3000))
        e.upload(frame)
        want = e.detect()
        assert len(want) > 100
        # same pairs, same values: map the upload indices to slots and sort like the download does
        w = want.copy()
        w["i"], w["j"] = slot[want["i"]], slot[want["j"]]
        w = w[np.lexsort((w["j"], w["i"]))]
        assert got.tobytes() == w.tobytes()
        assert s.engine.counts()["n_candidates"] == e.counts()["n_candidates"]


def test_trajectory_rings_and_predict_match_the_per_field_path():
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    from rcd_b200.host.ingest import VehiclePositionStream
    n = 1500
    frame = W.uniform_frame(n, 32, map_size=700.0, drone_fraction=0.3)
    ids = [f"v{k}" for k in range(n)]
    rng = np.random.default_rng(2)
    with VehiclePositionStream(2048, 1 << 20) as s, FrameEngine(2048, 1 << 20) as e:
        e.history_configure(100)
        f = {k: v.copy() for k, v in frame.items()}
        for step in range(4):
            t = 100.0 + 0.5 * step
            # every step a different third of the vehicles stays silent (their state and ring keep)
            talk = np.flatnonzero(rng.random(n) < (1.0 if step == 0 else 0.67))
            s.handle_messages(_messages(f, ids, t, talk))
            e.history_append(talk.astype(np.uint32), f["px"][talk].astype(np.float64), f["py"][talk].astype(np.float64),
                             f["pz"][talk].astype(np.float64), np.full(len(talk), t))
            if step == 0:
                cur = {k: v.copy() for k, v in f.items()}
            else:
                for k in cur:
                    cur[k][talk] = f[k][talk]
            f = W.advance(f, 0.5, rng)
        e.upload(cur)
        want_codes = e.history_classify()
        want = e.predict()
        got = s.detect_all_vehicles(predict=True)
        got_codes = s.engine.history_classify()
        assert np.array_equal(got_codes, want_codes) and len(set(want_codes.tolist())) >= 2
        assert len(want) > 50 and got.tobytes() == want.tobytes()


def test_several_messages_for_one_vehicle_in_one_batch():
    from rcd_b200.host import workloads as W
    from rcd_b200.host.ingest import VehiclePositionStream
    n = 400
    f0 = W.uniform_frame(n, 33, map_size=300.0)
    rng = np.random.default_rng(3)
    f1 = W.advance({k: v.copy() for k, v in f0.items()}, 0.5, rng)
    f2 = W.advance({k: v.copy() for k, v in f1.items()}, 0.5, rng)
    ids = [f"v{k}" for k in range(n)]
    one = "\n".join([_messages(f0, ids, 1.0), _messages(f1, ids, 1.5), _messages(f2, ids, 2.0)])
    with VehiclePositionStream(512, 1 << 18) as a, VehiclePositionStream(512, 1 << 18) as b:
        assert a.handle_messages(one) == 3 * n           # one batch, seq 0..2
        for f, t in ((f0, 1.0), (f1, 1.5), (f2, 2.0)):   # three batches
            b.handle_messages(_messages(f, ids, t))
        pa, pb = a.detect_all_vehicles(), b.detect_all_vehicles()
        assert np.array_equal(a.engine.history_classify(), b.engine.history_classify())
        assert pa.tobytes() == pb.tobytes()
        assert a.detect_all_vehicles(predict=False).tobytes() == b.detect_all_vehicles(predict=False).tobytes()
