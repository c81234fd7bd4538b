"""Batched ingest decode (csrc/rcd_ingest.hpp) against what the reference does
with the same message texts: EarlyWarningSystem._handle_vehicle_position (warning_system.py:638-678),
recorded in tests/golden/ingest_messages.json by tests/golden/make_golden.py (reference bytecode under
the shim).  The decoder is host code inside the CUDA library; these tests need no GPU."""
import json
import math
import os
import random

import numpy as np
import pytest

from tests import ingest_cases as C

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ingest_messages.json")


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _same(a: float, b: float) -> bool:
    return (a == b and math.copysign(1.0, a) == math.copysign(1.0, b)) or (a != a and b != b)


def _check_record(g, rec, want):
    """want = [id, 11 float reprs (pos, vel, acc, heading, size), type, timestamp repr]"""
    vals = [float(x) for x in want[1:12]]
    assert g.id_of(int(rec["slot"])) == want[0]
    assert g.type_of(int(rec["type"])) == want[12]
    for k, name in enumerate(("x", "y", "z")):
        assert _same(float(rec[name]), vals[k]), (name, rec[name], vals[k])  # float64, bit for bit
    assert _same(float(rec["timestamp"]), float(want[13]))
    with np.errstate(over="ignore"):
        for k, name in enumerate(("vx", "vy", "vz", "ax", "ay", "az")):
            assert _same(float(rec[name]), float(np.float32(vals[3 + k]))), name
        assert _same(float(rec["heading"]), float(np.float32(vals[9])))
        assert _same(float(rec["size"]), float(np.float32(vals[10])))


def test_decode_matches_reference_handler():
    from rcd_b200.host.ingest import VehicleIngest
    rows = _golden()
    with VehicleIngest(threads=1) as g:
        rec, max_seq = g.decode("\n".join(r["text"] for r in rows))
        kept = [r for r in rows if r["vehicle"] is not None]
        assert len(rec) == len(kept)
        assert g.bad_messages == len(rows) - len(kept)  # dropped exactly where the reference drops
        seen = {}
        for r, row in zip(rec, kept):
            _check_record(g, r, row["vehicle"])
            vid = row["vehicle"][0]
            assert int(r["seq"]) == seen.get(vid, 0)  # occurrence number inside the batch
            seen[vid] = seen.get(vid, 0) + 1
        assert max_seq == max(seen.values()) - 1
        assert g.n_objects == len(seen)
        # slots are dense and in order of first appearance
        first = list(dict.fromkeys(row["vehicle"][0] for row in kept))
        assert g.ids() == first


def test_oracle_parse_is_pinned_to_the_reference():
    from oracle import oracle as O
    for row in _golden():
        got = O.parse_vehicle_message(row["text"])
        if row["vehicle"] is None:
            assert got is None, row["text"]
            continue
        w = row["vehicle"]
        flat = [got["id"], *got["position"], *got["velocity"], *got["acceleration"], got["heading"], got["size"],
                got["type"], got["timestamp"]]
        assert flat[0] == w[0] and flat[12] == w[12]
        for a, b in zip(flat[1:12] + [flat[13]], w[1:12] + [w[13]]):
            assert _same(float(a), float(b))


def test_one_message_at_a_time_and_persistent_slots():
    from rcd_b200.host.ingest import VehicleIngest
    msgs = C.seeded_messages(50, 5)
    with VehicleIngest(threads=1) as g:
        for k, m in enumerate(msgs):
            rec, max_seq = g.decode(m)
            assert len(rec) == 1 and max_seq == 0 and int(rec[0]["slot"]) == k and int(rec[0]["seq"]) == 0
        rec, _ = g.decode(msgs[7])  # a known id keeps its slot in later batches
        assert int(rec[0]["slot"]) == 7 and int(rec[0]["seq"]) == 0
        assert g.slot_of("vehicle-7") == 7 and g.slot_of("nobody") is None
        assert g.n_objects == 50 and g.bad_messages == 0


def test_array_wrapper_and_separators():
    from rcd_b200.host.ingest import VehicleIngest
    msgs = C.seeded_messages(20, 6)
    with VehicleIngest(threads=1) as a, VehicleIngest(threads=1) as b, VehicleIngest(threads=1) as c:
        ra, _ = a.decode("\n".join(msgs))
        rb, _ = b.decode("[" + ",\n ".join(msgs) + "]")
        rc, _ = c.decode(" ".join(msgs).encode("utf-8"))
        assert ra.tobytes() == rb.tobytes() == rc.tobytes()
        assert a.bad_messages == b.bad_messages == c.bad_messages == 0


def test_parallel_decode_equals_sequential():
    from rcd_b200.host.ingest import VehicleIngest
    rng = random.Random(9)
    msgs = []
    for k in range(6000):  # > 1 MiB, with repeated ids and some broken lines
        msgs.append(C.reference_message(rng, k, 1000.0 + k, vid=f"vehicle-{rng.randrange(4000)}"))
        if k % 500 == 17:
            msgs.append('{"id": "cut", "position": {"x": 1')
    text = "\n".join(msgs)
    assert len(text) > (1 << 20)
    with VehicleIngest(threads=1) as a, VehicleIngest(threads=5) as b:
        ra, sa = a.decode(text)
        rb, sb = b.decode(text)
        assert len(ra) == 6000 and ra.tobytes() == rb.tobytes() and sa == sb and sa >= 1
        assert a.bad_messages == b.bad_messages == 12
        assert a.ids() == b.ids()


def test_record_buffer_too_small_is_an_error():
    from rcd_b200.host import _native as N
    from rcd_b200.host.ingest import VehicleIngest
    msgs = C.seeded_messages(10, 8)
    with VehicleIngest(threads=1) as g:
        out = np.empty(4, dtype=N.RECORD_DTYPE)
        with pytest.raises(N.NativeError) as e:
            g.decode("\n".join(msgs), out=out)
        assert e.value.code == N.RCD_ECAPACITY


def test_record_buffer_too_small_changes_no_state():
    """ECAPACITY is reported before ids are interned or sequence numbers advance: the retry with a larger buffer
    returns what a first call would have returned."""
    from rcd_b200.host import _native as N
    from rcd_b200.host.ingest import VehicleIngest
    text = "\n".join(C.seeded_messages(40, 11))
    with VehicleIngest(threads=1) as g, VehicleIngest(threads=1) as fresh:
        with pytest.raises(N.NativeError):
            g.decode(text, out=np.empty(4, dtype=N.RECORD_DTYPE))
        assert g.n_objects == 0
        ra, sa = g.decode(text)
        rb, sb = fresh.decode(text)
        assert ra.tobytes() == rb.tobytes() and sa == sb and g.ids() == fresh.ids()


def test_one_misbehaving_producer_does_not_discard_the_batch():
    """Limits of the record format degrade per message (the reference handles every message on its own,
    warning_system.py:638-678): the 256th message of one vehicle in a call is dropped, type strings beyond 255
    share a code, and a full id table rejects new vehicles while known ones keep being served."""
    from rcd_b200.host.ingest import VehicleIngest
    rng = random.Random(3)
    good = [C.reference_message(rng, k, 10.0 + k, vid=f"car-{k}") for k in range(20)]
    flood = [C.reference_message(rng, 100 + k, 50.0 + k, vid="chatty") for k in range(300)]
    with VehicleIngest(threads=1) as g:
        rec, max_seq = g.decode("\n".join(good[:10] + flood + good[10:]))
        assert len(rec) == 20 + 255 and max_seq == 254
        assert g.bad_messages == 45 and g.rejected == 45
        assert sorted(set(rec["slot"].tolist())) == list(range(21))  # every other vehicle was applied
        # the next call starts new sequence numbers: the chatty vehicle is served again
        rec2, _ = g.decode(flood[0])
        assert len(rec2) == 1 and rec2["seq"][0] == 0
    with VehicleIngest(threads=1) as g:
        msgs = []
        for k in range(300):
            m = json.loads(C.reference_message(rng, k, 1.0 + k, vid=f"v{k}"))
            m["type"] = f"kind-{k}"
            msgs.append(json.dumps(m))
        rec, _ = g.decode("\n".join(msgs))
        assert len(rec) == 300 and g.bad_messages == 0 and g.n_types == 255
        assert rec["type"][:255].tolist() == list(range(255)) and set(rec["type"][255:].tolist()) == {255}
    with VehicleIngest(threads=1) as g:
        g.set_limit(5)
        rec, _ = g.decode("\n".join(good))
        assert len(rec) == 5 and g.n_objects == 5 and g.rejected == 15
        rec, _ = g.decode("\n".join(good))  # known vehicles keep being served, forever
        assert len(rec) == 5 and rec["slot"].tolist() == list(range(5)) and g.rejected == 30


@pytest.mark.needs_reference
def test_fresh_messages_against_the_live_reference():
    from oracle import ref_shim as S
    from rcd_b200.host.ingest import VehicleIngest
    msgs = C.seeded_messages(300, 4242)
    want = S.run_handle_position_A(msgs)
    with VehicleIngest(threads=1) as g:
        rec, _ = g.decode("\n".join(msgs))
        assert len(rec) == len(want) == 300
        for r, w in zip(rec, want):
            _check_record(g, r, [w[0]] + [repr(float(x)) for x in w[1:12]] + [w[12], repr(float(w[13]))])


def test_mutated_messages_never_disagree_with_the_python_path():
    """Seeded fuzz: random deletions / insertions / replacements in valid messages.  The decoder must accept
    exactly the texts the reference path accepts (json.loads + the handler's field reads), with identical
    values -- except that a complete message followed by garbage is one message plus one dropped message
    for a stream decoder, where json.loads rejects the whole text ("Extra data")."""
    from oracle import oracle as O
    from rcd_b200.host.ingest import VehicleIngest
    rng = random.Random(20261018)
    base = C.seeded_messages(120, 3) + [t for t, _ in C.edge_messages()]
    alphabet = list('{}[]":,.-+eE0123456789 \\ntrue falsnNaInity\t') + ['\\u00e9', '\\ud83d', '"x"', '"y"', '"id"']

    def mutate(s):
        s = list(s)
        for _ in range(rng.choice([1, 1, 2, 3, 6])):
            op, pos = rng.random(), rng.randrange(len(s) + 1)
            if op < 0.35 and s:
                del s[min(pos, len(s) - 1)]
            elif op < 0.7:
                s.insert(pos, rng.choice(alphabet))
            elif s:
                s[min(pos, len(s) - 1)] = rng.choice(alphabet)
        return "".join(s).replace("\n", " ")

    g = VehicleIngest(threads=1)
    accepted = dropped = 0
    for k in range(6000):
        if k % 100 == 0:  # fresh id / type tables: the mutations invent type strings (limit 255 per table)
            g.close()
            g = VehicleIngest(threads=1)
        m = mutate(rng.choice(base))
        want = O.parse_vehicle_message(m)
        rec, _ = g.decode(m)
        if want is None:
            try:
                json.loads(m)
                extra = False
            except json.JSONDecodeError as e:
                extra = e.msg == "Extra data"
            except RecursionError:
                continue
            assert len(rec) == 0 or extra, m
            dropped += 1
            continue
        assert len(rec) == 1, m
        vals = [float(v) for v in (*want["position"], *want["velocity"], *want["acceleration"], want["heading"], want["size"])]
        _check_record(g, rec[0], [want["id"]] + [repr(v) for v in vals] + [want["type"], repr(float(want["timestamp"]))])
        accepted += 1
    g.close()
    assert accepted > 500 and dropped > 2000
