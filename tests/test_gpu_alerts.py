"""Device alert table (csrc/rcd_alerts.cuh) against the oracle's AlertTable, which is pinned to the
reference's AlertManager (tests/test_alerts_cpu.py)."""
import numpy as np
import pytest

from tests import alert_cases as A

pytestmark = pytest.mark.gpu


def _pairs(op_risks):
    from oracle import oracle as O
    from rcd_b200.host import _native as N
    p = np.zeros(len(op_risks), dtype=N.PAIR_DTYPE)
    for k, (v, o, risk, ttc, dist) in enumerate(op_risks):
        p[k]["i"], p[k]["j"] = int(v[1:]), int(o[1:])
        p[k]["risk"], p[k]["ttc"], p[k]["distance"] = risk, ttc, dist
        # the priority travels with the pair, as the frame kernels emit it (-1 below RISK_LEVEL_LOW)
        p[k]["priority"] = O.alert_priority(risk, ttc) if risk >= 0.3 else -1
    return p


def _events(ev):
    from rcd_b200.host import _native as N
    kind = {N.ALERT_CREATED: "created", N.ALERT_PRIORITY_CHANGED: "changed", N.ALERT_REFRESHED: "refreshed", N.ALERT_EXPIRED: "expired"}
    return sorted((kind[int(e["kind"])], f"v{int(e['i'])}", f"v{int(e['j'])}",
                   -1 if int(e["kind"]) == N.ALERT_EXPIRED else int(e["priority"]),
                   -1 if int(e["kind"]) == N.ALERT_EXPIRED else int(e["old_priority"])) for e in ev)


def _table(dl):
    return sorted((f"v{int(e['i'])}", f"v{int(e['j'])}", float(e["risk"]), float(e["ttc"]), int(e["priority"]),
                   float(e["timestamp"]), bool(e["acknowledged"])) for e in dl)


def test_scenario_matches_the_oracle_op_by_op():
    from rcd_b200.host.engine import FrameEngine
    from tests.test_alerts_cpu import run_oracle
    script = A.scenario()
    want = run_oracle(script)
    with FrameEngine(64, 64) as e:
        e.alerts_configure(4096)
        live = 0
        for op, (w_events, w_table) in zip(script, want):
            if op[0] == "process":
                ev, st = e.alerts_update_pairs(_pairs(op[2]), op[1], report_refreshed=True)
                assert _events(ev) == w_events
                assert st["n_created"] == sum(x[0] == "created" for x in w_events)
                assert st["n_changed"] == sum(x[0] == "changed" for x in w_events)
                assert st["n_refreshed"] == sum(x[0] == "refreshed" for x in w_events)
                ids = ev["alert_id"][ev["kind"] == 1]
                assert len(set(ids.tolist())) == len(ids)  # every created alert gets its own number
            elif op[0] == "ack":
                e.alerts_acknowledge([int(v[1:]) for v, _ in op[1]], [int(o[1:]) for _, o in op[1]])
                st = None
            else:
                ev, st = e.alerts_expire(op[1], 30.0)
                assert _events(ev) == w_events and st["n_expired"] == len(w_events)
            dl = e.alerts_download()
            assert _table(dl) == w_table  # risk / ttc / timestamp bit for bit, priority, acknowledged
            if st is not None:
                assert st["n_live"] == len(w_table) and st["n_dropped"] == 0
        assert len(e.alerts_download()) == 0


def test_only_changes_are_reported_by_default_and_frames_feed_the_table_on_the_device():
    """rcd_alerts_update folds the pairs of the last frame where they lie: same result as downloading them
    and passing them back."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    frame = W.uniform_frame(4000, 41, map_size=500.0, drone_fraction=0.3)
    with FrameEngine(4096, 1 << 20) as a, FrameEngine(4096, 1 << 20) as b:
        for e in (a, b):
            e.alerts_configure(1 << 18)
            e.upload(frame)
            e.step(N.MODE_DETECT)
            e.step(N.MODE_PREDICT, append=True)  # detect + predict risks for the same (i, j): predicted wins
        pairs = a.download()
        n_alerting = int((pairs["priority"] >= 0).sum())
        keys = set(zip(pairs["i"][pairs["priority"] >= 0].tolist(), pairs["j"][pairs["priority"] >= 0].tolist()))
        assert n_alerting > len(keys) > 1000
        ev_a, st_a = a.alerts_update(10.0)
        ev_b, st_b = b.alerts_update_pairs(pairs, 10.0)
        assert st_a == st_b and st_a["n_created"] == len(keys) and st_a["n_live"] == len(keys)
        assert st_a["n_created"] + st_a["n_changed"] + st_a["n_refreshed"] == n_alerting
        assert len(ev_a) == st_a["n_created"] + st_a["n_changed"]  # refreshed alerts are not reported
        ta, tb = a.alerts_download(), b.alerts_download()
        drop = lambda t: [(int(x["i"]), int(x["j"]), float(x["risk"]), float(x["ttc"]), int(x["priority"])) for x in t]
        assert drop(ta) == drop(tb)
        # the later (predicted) risk of a pair is the one the table keeps
        last = {}
        for p in pairs[pairs["priority"] >= 0]:
            k = (int(p["i"]), int(p["j"]))
            if k not in last or p["predicted"] >= last[k]["predicted"]:
                last[k] = p
        assert drop(ta) == sorted((k[0], k[1], float(p["risk"]), float(p["ttc"]), int(p["priority"])) for k, p in last.items())
        # same frame again, 5 s later: nothing is created, nothing changes priority
        ev2, st2 = a.alerts_update(15.0)
        assert st2["n_created"] == 0 and st2["n_live"] == len(keys) and len(ev2) == st2["n_changed"]
        # a table that is too small reports what it dropped instead of corrupting anything
        b.alerts_configure(256)
        _ev, st3 = b.alerts_update(20.0)
        assert st3["n_dropped"] > 0 and st3["n_live"] <= 1024 and st3["n_live"] + st3["n_dropped"] >= len(keys)


def test_messages_to_alerts_end_to_end():
    """The reference's loop (warning_system.py:638-714) in batch form: message buffer -> frame ->
    predict for every vehicle -> alerts; only changes surface in Python."""
    import json
    from rcd_b200.host import workloads as W
    from rcd_b200.host.ingest import VehiclePositionStream
    from rcd_b200.host.warning_system import AlertManager
    n = 1200
    f = W.uniform_frame(n, 51, map_size=400.0)
    rng = np.random.default_rng(5)
    msg = lambda fr, t: "\n".join(json.dumps({
        "id": f"car-{i}", "position": {"x": float(fr["px"][i]), "y": float(fr["py"][i]), "z": float(fr["pz"][i])},
        "velocity": {"x": float(fr["vx"][i]), "y": float(fr["vy"][i]), "z": float(fr["vz"][i])},
        "acceleration": {"x": float(fr["ax"][i]), "y": float(fr["ay"][i]), "z": float(fr["az"][i])},
        "heading": float(fr["heading"][i]), "size": float(fr["size"][i]), "type": "car", "timestamp": t}) for i in range(n))
    with VehiclePositionStream(2048, 1 << 20) as s:
        am = AlertManager()
        am.attach_engine(s.engine, s.ingest.id_of, 1 << 18)
        seen = {}
        for step in range(3):
            s.handle_messages(msg(f, 100.0 + 0.5 * step))
            pairs = s.detect_all_vehicles(predict=True)
            alerts = am.process_frame(now=100.0 + 0.5 * step)
            alerting = pairs[pairs["priority"] >= 0]
            keys = {(s.ingest.id_of(int(p["i"])), s.ingest.id_of(int(p["j"]))): int(p["priority"]) for p in alerting}
            want = {k for k, pr in keys.items() if seen.get(k) != pr}  # new, or priority differs from last time
            assert {(a.vehicle_id, a.other_vehicle_id) for a in alerts} == want
            assert all(a.priority == keys[(a.vehicle_id, a.other_vehicle_id)] for a in alerts)
            seen.update(keys)
            assert am.get_stats()["active_alerts"] == len(seen)
            f = W.advance(f, 0.5, rng, map_size=(400.0, 400.0))
        assert len(seen) > 100
        gone = am.cleanup_expired(now=200.0)
        assert set(gone) == set(seen) and am.get_stats()["active_alerts"] == 0


def test_the_same_risk_twice_in_one_pass_is_create_then_silent_refresh():
    """A vehicle without history owes every risk twice (detect_collisions + the fall-back of predict_collisions,
    collision_detection.py:590-592): the copies sit next to each other in the pair buffer.  The reference's loop
    creates the alert with the first copy and refreshes it with the second (warning_system.py:259-285) -- never a
    priority change, never a half-written entry."""
    from rcd_b200.host import _native as N
    from rcd_b200.host.engine import FrameEngine
    rng = np.random.default_rng(5)
    n = 4000
    base = np.zeros(n, dtype=N.PAIR_DTYPE)
    base["i"] = rng.permutation(50_000)[:n].astype(np.uint32)
    base["j"] = base["i"] + 1 + rng.integers(0, 1000, n).astype(np.uint32)
    base["risk"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    base["ttc"] = rng.uniform(0.0, 10.0, n).astype(np.float32)
    base["priority"] = rng.integers(0, 4, n).astype(np.int8)
    twice = np.repeat(base, 2)  # adjacent copies: same warp (31 of 32 times) or neighbouring warps
    with FrameEngine(64, 64) as e:
        e.alerts_configure(16384)
        ev, st = e.alerts_update_pairs(twice, 100.0, report_refreshed=True)
        assert st["n_created"] == n and st["n_refreshed"] == n and st["n_changed"] == 0 and st["n_live"] == n
        created = ev[ev["kind"] == N.ALERT_CREATED]
        refreshed = ev[ev["kind"] == N.ALERT_REFRESHED]
        assert len(created) == n and len(refreshed) == n and len(ev) == 2 * n
        ids = {(int(r["i"]), int(r["j"])): int(r["alert_id"]) for r in created}
        assert len(set(ids.values())) == n
        for r in refreshed:  # the refresh reports the number and the priority the creation gave the alert
            assert int(r["alert_id"]) == ids[(int(r["i"]), int(r["j"]))]
            assert int(r["old_priority"]) == int(r["priority"])
        # the next frame sees complete entries
        ev, st = e.alerts_update_pairs(base, 101.0)
        assert st["n_refreshed"] == n and st["n_created"] == 0 and st["n_changed"] == 0 and len(ev) == 0


def test_summary_delivery_equals_the_synchronous_calls_frame_by_frame():
    """rcd_summary_begin / _finish: frame k's alert changes, per-object risk counts and totals arrive complete
    although frame k + 1 was uploaded and stepped in between; they equal rcd_alerts_update + the pair list of a
    second engine that runs the same frames one at a time.  Compact pair records equal the narrowed full records."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 6000
    rng = np.random.default_rng(3)
    frames = [W.uniform_frame(n, 201, map_size=1000.0, drone_fraction=0.3)]
    for _ in range(4):
        frames.append(W.advance(frames[-1], 0.3, rng, map_size=(1000.0, 1000.0)))
    pat = W.random_patterns(n, 202)
    key = lambda ev: sorted((int(e["i"]), int(e["j"]), int(e["kind"]), int(e["priority"]), int(e["old_priority"]),
                             float(e["risk"]), float(e["ttc"]), float(e["distance"]), float(e["timestamp"])) for e in ev)
    with FrameEngine(n, 1 << 20) as ref, FrameEngine(n, 1 << 20) as e:
        ref.alerts_configure(1 << 18)
        e.alerts_configure(1 << 18)
        want, got = [], []
        def step(eng, k):  # fused frames and detect + appended predict frames alternate (the fold's detect pass is bounded
            if k % 2:      # by where the records with predicted = 0 end: both layouts of the pair buffer are exercised)
                eng.step(N.MODE_DETECT)
                eng.step(N.MODE_PREDICT, append=True)
            else:
                eng.step(N.MODE_PREDICT, with_detect=True)

        for k, f in enumerate(frames):
            ref.upload(f); ref.set_patterns(pat); step(ref, k)
            pairs = ref.download()
            ev, st = ref.alerts_update(100.0 + k)
            assert np.array_equal(ref.risk_counts(), np.bincount(pairs["i"], minlength=n))
            want.append((key(ev), st, np.bincount(pairs["i"], minlength=n).astype(np.uint32), ref.counts()))
            e.upload(f); e.set_patterns(pat); step(e, k)
            if k:
                got.append(e.summary_finish(risk_counts=np.zeros(n, np.uint32)))
            e.summary_begin(100.0 + k)
            with pytest.raises(N.NativeError):
                e.download_begin(np.zeros(16, N.PAIR_DTYPE))  # one delivery at a time
        got.append(e.summary_finish(risk_counts=np.zeros(n, np.uint32)))
        assert sum(len(w[0]) for w in want) > 1000
        for (w_ev, w_st, w_rc, w_c), (g_ev, g_st, g_rc, g_c) in zip(want, got):
            assert key(g_ev) == w_ev
            assert g_st == w_st
            assert np.array_equal(g_rc, w_rc)
            for name in ("n_pairs", "n_candidates", "n_potential", "n_high_risk", "n_alerts", "n_objects"):
                assert g_c[name] == w_c[name], name
        # compact records
        full = ref.download()
        buf = np.zeros(len(full) + 8, dtype=N.PAIR_COMPACT_DTYPE)
        ref.download_begin_compact(buf)
        compact, counts = ref.download_finish()
        assert counts["n_pairs"] == len(full) == len(compact)
        compact = np.sort(compact, order=["i", "j", "predicted"], kind="stable")
        for name in ("i", "j", "ttc", "distance", "rel_speed", "risk", "t_closest", "priority", "offset", "predicted"):
            assert np.array_equal(compact[name], full[name]), name
