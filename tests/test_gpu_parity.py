"""GPU parity: the CUDA path through the C-ABI against (a) the golden vectors produced by the
reference's own bytecode and (b) the CPU oracle on fresh seeded frames.  Pair sets bit-exact,
values within tolerance (tests/gpu_helpers.py)."""
import numpy as np
import pytest

from tests.gpu_helpers import compare_counts, compare_pairs
from tests.helpers import assert_pairs_equal, f64_frame, load_golden

pytestmark = pytest.mark.gpu

DETECT = ["detect_dense3d.npz", "detect_2d_noaccel.npz", "detect_r60_t4.npz", "detect_city1k.npz"]
PREDICT = ["predict_dense3d.npz", "predict_2d_allcv.npz", "predict_city300.npz"]


@pytest.fixture(scope="module")
def eng():
    from rcd_b200.host.engine import FrameEngine
    e = FrameEngine(max_objects=1_100_000, max_pairs=4_000_000, profile=True, count_predict_candidates=True)
    yield e
    e.close()


def _oracle():
    from oracle import oracle as O
    return O


def _risks_from_golden(tab, with_offset=False):
    from oracle.oracle import RISK_DTYPE
    r = np.zeros(len(tab), RISK_DTYPE)
    for k, c in enumerate(("i", "j", "ttc", "distance", "rel_speed", "risk", "cx", "cy", "cz")):
        r[c] = tab[:, k]
    return r


@pytest.mark.parametrize("name", DETECT)
def test_detect_golden(eng, name):
    z, frame = load_golden(name)
    eng.upload(frame)
    got = eng.detect(float(z["R"]), float(z["T"]))
    want = z["risks"]
    assert_pairs_equal(np.stack([got["i"], got["j"]], 1), want[:, :2], "detect pair set vs reference golden")
    ora = _oracle().frame_A(f64_frame(frame), "detect", R=float(z["R"]), T=float(z["T"]))
    compare_pairs(got, ora["risks"], "detect")
    c = eng.counts()
    assert c["n_candidates"] == len(z["candidates"])
    assert np.array_equal(eng.candidate_counts(), np.bincount(z["candidates"][:, 0], minlength=len(frame["px"])))
    assert c["n_potential"] == int(z["stat_potential"]) and c["n_high_risk"] == int(z["stat_high"])
    # stage-2 values ride along with every emitted pair
    pots = {(int(p[0]), int(p[1])): (p[2], p[3]) for p in z["potentials"]}
    for r in got:
        tc, cd = pots[(int(r["i"]), int(r["j"]))]
        assert abs(r["t_closest"] - tc) <= 1e-6 * abs(tc) + 1e-7 and abs(r["d_closest"] - cd) <= 1e-6 * cd + 1e-7


@pytest.mark.parametrize("name", PREDICT)
def test_predict_golden(eng, name):
    z, frame = load_golden(name)
    eng.upload(frame)
    eng.set_patterns(z["pattern"])
    got = eng.predict()
    assert_pairs_equal(np.stack([got["i"], got["j"]], 1), z["risks"][:, :2], "predict pair set vs reference golden")
    assert np.array_equal(got["predicted"].astype(bool), z["is_predicted"])
    ora = _oracle().frame_A(f64_frame(frame), "predict", pattern_codes=z["pattern"])
    compare_pairs(got, ora["risks"], "predict")
    compare_counts(eng.counts(), eng.candidate_counts(), ora, "predict")


def test_compute_node_golden(eng):
    z, frame = load_golden("implB_dense.npz")
    eng.upload(frame)
    eng.set_patterns(z["has_history"].astype(np.uint8))
    got = eng.compute_node(100.0)
    assert_pairs_equal(np.stack([got["i"], got["j"]], 1), z["risks"][:, :2], "compute-node pair set vs golden")
    ora = _oracle().frame_B(f64_frame(frame), z["has_history"])
    compare_pairs(got, ora["risks"], "compute_node")
    compare_counts(eng.counts(), eng.candidate_counts(), ora, "compute_node")


@pytest.mark.parametrize("seed,n,box,drones,static_bounds", [
    (11, 3000, 900.0, 0.3, False), (12, 6000, 1500.0, 0.0, True), (13, 2049, 500.0, 0.5, False),
    (14, 20000, 4000.0, 0.3, True), (15, 4097, 700.0, 0.2, False)])
def test_detect_and_predict_vs_oracle(seed, n, box, drones, static_bounds):
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    O = _oracle()
    frame = W.uniform_frame(n, seed, map_size=box, drone_fraction=drones)
    pat = W.random_patterns(n, seed + 100)
    bounds = ((0, 0, 0), (box, box, 100.0)) if static_bounds else None
    count = bool(seed % 2)  # exercise both predict kernels (with / without candidate counting)
    with FrameEngine(n, 64 * n, world_bounds=bounds, count_predict_candidates=count) as e:
        e.upload(frame)
        got = e.detect()
        ora = O.frame_A(f64_frame(frame), "detect")
        compare_pairs(got, ora["risks"], "detect")
        compare_counts(e.counts(), e.candidate_counts(), ora, "detect")
        e.set_patterns(pat)
        got = e.predict()
        ora = O.frame_A(f64_frame(frame), "predict", pattern_codes=pat)
        compare_pairs(got, ora["risks"], "predict")
        compare_counts(e.counts(), e.candidate_counts(), ora, "predict", candidates=count)


def test_append_mode_is_detect_plus_predict(eng):
    from rcd_b200.host import _native as N
    from rcd_b200.host import workloads as W
    frame = W.uniform_frame(5000, 21, map_size=1200.0, drone_fraction=0.3)
    eng.upload(frame)
    eng.set_patterns(np.full(5000, 1, np.uint8))
    d = eng.detect()
    cd = eng.counts()
    p = eng.predict()
    cp = eng.counts()
    eng.step(N.MODE_DETECT)
    eng.step(N.MODE_PREDICT, append=True)
    both = eng.download()
    cb = eng.counts()
    assert len(both) == len(d) + len(p) > 0
    assert cb["n_candidates"] == cd["n_candidates"] + cp["n_candidates"]
    assert np.array_equal(np.sort(both[both["predicted"] == 0], order=["i", "j"]), d)
    assert np.array_equal(np.sort(both[both["predicted"] == 1], order=["i", "j"]), p)


@pytest.mark.parametrize("R,T", [(10.0, 10.0), (250.0, 3.0), (100.0, 0.05), (37.5, 20.0)])
def test_detect_nondefault_radius_and_window(eng, R, T):
    from rcd_b200.host import workloads as W
    frame = W.uniform_frame(4000, 31, map_size=1000.0, drone_fraction=0.3)
    eng.upload(frame)
    got = eng.detect(R, T)
    ora = _oracle().frame_A(f64_frame(frame), "detect", R=R, T=T)
    compare_pairs(got, ora["risks"], "detect")
    compare_counts(eng.counts(), eng.candidate_counts(), ora, "detect")


def test_edge_cases(eng):
    from rcd_b200.host import workloads as W
    O = _oracle()
    # empty frame and single object
    for n in (0, 1):
        eng.upload(W.uniform_frame(n, 1))
        assert len(eng.detect()) == 0 and len(eng.predict()) == 0
        c = eng.counts()
        assert c["n_candidates"] == 0 and c["n_pairs"] == 0 and c["n_objects"] == n
    # coincident objects, negative coordinates, everything in one cell, stationary vehicles
    f = W.uniform_frame(300, 41, map_size=40.0, drone_fraction=0.3)
    for k in ("px", "py"):
        f[k] = (f[k] - 20.0).astype(np.float32)
    f["px"][:10] = f["px"][0]; f["py"][:10] = f["py"][0]; f["pz"][:10] = f["pz"][0]
    for k in ("vx", "vy", "vz", "ax", "ay", "az"):
        f[k][100:160] = 0.0
    pat = W.random_patterns(300, 42)
    eng.upload(f)
    got = eng.detect()
    ora = O.frame_A(f64_frame(f), "detect")
    compare_pairs(got, ora["risks"], "detect")
    compare_counts(eng.counts(), eng.candidate_counts(), ora, "detect")
    eng.set_patterns(pat)
    got = eng.predict()
    ora = O.frame_A(f64_frame(f), "predict", pattern_codes=pat)
    compare_pairs(got, ora["risks"], "predict")
    compare_counts(eng.counts(), eng.candidate_counts(), ora, "predict")


def test_objects_outside_static_bounds_are_clamped_not_lost():
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    frame = W.uniform_frame(5000, 51, map_size=1500.0, drone_fraction=0.3)
    # the grid only covers the middle of the map: everything else lands in border cells
    with FrameEngine(5000, 500000, world_bounds=((500, 500, 20), (1000, 1000, 60))) as e:
        e.upload(frame)
        got = e.detect()
        ora = _oracle().frame_A(f64_frame(frame), "detect")
        compare_pairs(got, ora["risks"], "detect")
        compare_counts(e.counts(), e.candidate_counts(), ora, "detect")


def test_pair_buffer_overflow_keeps_counts_exact():
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    frame = W.uniform_frame(3000, 61, map_size=300.0, drone_fraction=0.3)
    ora = _oracle().frame_A(f64_frame(frame), "detect")
    assert ora["counts"][2] > 64
    with FrameEngine(3000, 64) as e:
        e.upload(frame)
        got = e.detect()
        c = e.counts()
        assert len(got) == 64 and c["n_written"] == 64 and c["n_pairs"] == int(ora["counts"][2])


def test_caller_ids_are_reported(eng):
    from rcd_b200.host import workloads as W
    frame = W.uniform_frame(2000, 71, map_size=400.0)
    ids = (np.arange(2000, dtype=np.uint32) * 7 + 1000)
    eng.upload(frame, ids=ids)
    got = eng.detect()
    ora = _oracle().frame_A(f64_frame(frame), "detect")["risks"]
    assert len(got) == len(ora) > 0
    assert_pairs_equal(np.stack([got["i"], got["j"]], 1), np.stack([ids[ora["i"]], ids[ora["j"]]], 1), "ids")


def test_radius_queries(eng):
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "scalars.npz"))
    frame = {k[len("near_frame_"):]: z[k] for k in z.files if k.startswith("near_frame_")}
    eng.upload(frame)
    res = eng.query_radius(z["near_q"], float(z["near_radius"]))
    off = z["near_off"]
    for k, ids in enumerate(res):
        assert np.array_equal(ids, z["near_ids"][off[k]:off[k + 1]]), f"query {k}"
    # after a frame was built with a different cell size, and with a radius larger than the cell
    eng.detect(30.0, 10.0)
    q = np.random.default_rng(5).uniform(-100, 800, (200, 3)).astype(np.float32)
    q[:, 2] *= 0.1
    for radius in (5.0, 120.0):
        res = eng.query_radius(q, radius)
        want = _oracle().query_radius(f64_frame(frame), q.astype(np.float64), radius)
        for a, b in zip(res, want):
            assert np.array_equal(a, b)


def test_pattern_classifier(eng):
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "scalars.npz"))
    off = z["hist_off"]
    n = len(off) - 1
    stride = int(np.max(np.diff(off)))
    samples = np.zeros((n, max(stride, 1), 4))
    count = np.diff(off).astype(np.uint32)
    for k in range(n):
        h = z["hist"][off[k]:off[k + 1]]
        h = h[np.argsort(h[:, 3], kind="stable")]  # the reference sorts by timestamp (stable)
        samples[k, :len(h)] = h
    got = eng.classify_patterns(samples, count)
    assert np.array_equal(got, z["hist_class"].astype(np.uint8))


def test_two_slabs_with_halo_equal_one_domain():
    """Spatial slabs emulated on one GPU: each slab owns its objects, receives the
    packed halo of the other, and the union of the owners' results is the single-domain result."""
    import torch
    from rcd_b200.host import _native as N
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    n, box = 8000, 2000.0
    frame = W.uniform_frame(n, 81, map_size=box, drone_fraction=0.3)
    ids = np.arange(n, dtype=np.uint32)
    cut = float(np.median(frame["px"]))
    lo, hi = np.array([-1e9, cut], np.float32), np.array([cut, 1e9], np.float32)
    halo = 100.0 + 30.0 * 9.5 + 0.5 * 1.5 * 9.5 ** 2 + 1.0
    pat = np.full(n, 2, np.uint8)
    ora = _oracle().frame_A(f64_frame(frame), "predict", pattern_codes=pat)["risks"]
    engines, bufs, counts, owned = [], [], [], []
    for r in range(2):
        m = (frame["px"] >= lo[r]) & (frame["px"] < hi[r])
        owned.append(m)
        e = FrameEngine(n, 64 * n)
        e.upload(W.take(frame, m), ids=ids[m])
        buf = torch.empty((n, N.HALO_RECORD_WORDS), dtype=torch.int32, device="cuda")
        counts.append(e.halo_pack(lo, hi, r, halo, buf.data_ptr(), n))
        engines.append(e); bufs.append(buf)
    torch.cuda.synchronize()
    got = []
    for r in range(2):
        src = 1 - r
        assert counts[src][src] == 0 and counts[src][r] > 0
        engines[r].halo_append(bufs[src].data_ptr(), int(counts[src][r]))
        got.append(engines[r].predict())
        assert engines[r].counts()["n_owned"] == int(owned[r].sum())
    both = np.sort(np.concatenate(got), order=["i", "j"])
    compare_pairs(both, ora, "predict")
    for e in engines:
        e.close()


def test_full_size_100k_counts_and_pairs():
    """configs[2] at full size: 100k uniform 2-D vehicles; oracle runs in about a second."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    frame = W.make_workload("cfg3_100k_uniform2d")
    O = _oracle()
    with FrameEngine(100_000, 2_000_000, world_bounds=((0, 0, 0), (10000, 10000, 0))) as e:
        e.upload(frame)
        got = e.detect()
        ora = O.frame_A(f64_frame(frame), "detect", want_potentials=False)
        compare_pairs(got, ora["risks"], "detect")
        compare_counts(e.counts(), e.candidate_counts(), ora, "detect")
        pat = np.ones(100_000, np.uint8)
        e.set_patterns(pat)
        got = e.predict()
        ora = O.frame_A(f64_frame(frame), "predict", pattern_codes=pat, want_potentials=False)
        compare_pairs(got, ora["risks"], "predict")
        compare_counts(e.counts(), e.candidate_counts(), ora, "predict", candidates=False)


def test_full_size_1m_sampled_queries_and_symmetry():
    """configs[3] at full size (1 M mixed vehicles + drones, clustered, 3-D).  The oracle cannot run
    the whole frame in seconds, so: (a) every 211th object is queried by the oracle against the FULL
    index and must get exactly the GPU's pairs, detect and predict; (b) size-independent properties:
    the detect pair set is symmetric with bit-identical values (every term of the reference formula
    is symmetric) and per-object candidate counts sum to the frame total."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    O = _oracle()
    frame = W.make_workload("cfg4_1m_clustered3d")
    n = len(frame["px"])
    stride = 211
    f64 = f64_frame(frame)
    with FrameEngine(n, 24_000_000, world_bounds=((0, 0, 0), (31623, 31623, 100))) as e:
        e.upload(frame)
        got = e.detect()
        c = e.counts()
        assert c["n_pairs"] == len(got) and c["n_fallback"] == 0
        cand = e.candidate_counts()
        assert int(cand.astype(np.int64).sum()) == c["n_candidates"]
        # (b) symmetry
        fwd = np.stack([got["i"], got["j"]], 1).astype(np.int64)
        rev = fwd[:, ::-1]
        order = np.lexsort((rev[:, 1], rev[:, 0]))
        assert np.array_equal(fwd, rev[order])
        for col in ("ttc", "distance", "rel_speed", "risk", "cx", "cy", "cz", "priority"):
            assert np.array_equal(got[col], got[col][order]), col
        # (a) sampled queries, detect
        ora = O.frame_A(f64, "detect", want_potentials=False, query_stride=stride)
        sel = got[got["i"] % stride == 0]
        compare_pairs(sel, ora["risks"], "detect")
        q = np.arange(0, n, stride)
        assert np.array_equal(cand[q], ora["cand_count"][q])
        # (a) sampled queries, predict
        pat = W.random_patterns(n, 5)
        e.set_patterns(pat)
        got = e.predict()
        assert e.counts()["n_fallback"] == 0
        ora = O.frame_A(f64, "predict", pattern_codes=pat, want_potentials=False, query_stride=stride)
        compare_pairs(got[got["i"] % stride == 0], ora["risks"], "predict")


def test_permutation_invariance_and_idempotence():
    """Shuffling the upload order changes neither the pair set nor any value (ids travel with the
    objects); stepping the same frame twice gives identical output."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 50_000
    frame = W.hotspot_frame(n, 31, 6000.0, 5, radius_range=(400.0, 900.0))
    ids = np.arange(n, dtype=np.uint32)
    perm = np.random.default_rng(7).permutation(n)
    pat = W.random_patterns(n, 9)
    with FrameEngine(n, 8_000_000) as e:
        e.upload(frame, ids=ids)
        e.set_patterns(pat)
        a1 = e.predict()
        a2 = e.predict()
        assert np.array_equal(a1, a2)
        e.upload(W.take(frame, perm), ids=ids[perm])
        e.set_patterns(pat[perm])
        b = e.predict()
        assert np.array_equal(a1, b)
        d1 = e.detect()
        e.upload(frame, ids=ids)
        d2 = e.detect()
        assert np.array_equal(d1, d2) and len(d1) > 100


@pytest.mark.parametrize("case", ["mixed_patterns", "all_accelerating", "other_radius"])
def test_fused_detect_predict_equals_the_two_passes(case):
    """rcd_step(PREDICT | WITH_DETECT) must return exactly what step(DETECT) + step(PREDICT | APPEND) returns:
    same records (byte for byte after the sorted download), same totals."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 6000
    frame = W.uniform_frame(n, 61, map_size=1100.0, drone_fraction=0.3)
    pat = W.random_patterns(n, 9, p=(0.1, 0.3, 0.4, 0.2)) if case != "all_accelerating" else np.full(n, 2, np.uint8)
    R, T = (100.0, 10.0) if case != "other_radius" else (60.0, 4.0)
    with FrameEngine(n, 1 << 21) as a, FrameEngine(n, 1 << 21) as b:
        for e in (a, b):
            e.upload(frame)
            e.set_patterns(pat)
        a.step(N.MODE_DETECT, R, T)
        a.step(N.MODE_PREDICT, R, T, append=True)
        b.step(N.MODE_PREDICT, R, T, with_detect=True)
        pa, pb = a.download(), b.download()
        ca, cb = a.counts(), b.counts()
        assert len(pa) > 2000 and (pa["predicted"] == 0).sum() > 200 and (pa["predicted"] == 1).sum() > 200
        assert pa.tobytes() == pb.tobytes()
        for k in ("n_pairs", "n_candidates", "n_potential", "n_high_risk", "n_alerts", "n_objects", "n_owned"):
            assert ca[k] == cb[k], k
        assert cb["n_fallback"] == 0
        if case == "mixed_patterns":  # objects without history owe every risk twice (detect + predict fall-back)
            nohist = np.flatnonzero(pat == 3)
            twice = pb[np.isin(pb["i"], nohist)]
            assert len(twice) > 0 and len(twice) % 2 == 0
            assert twice[0::2].tobytes() == twice[1::2].tobytes()


def _fused_oracle(frame, pat):
    """What the fused frame must emit: detect for everyone + predict for everyone (objects without
    history fall back to detect there too: those risks are not `predicted`)."""
    O = _oracle()
    f64 = f64_frame(frame)
    d = O.frame_A(f64, "detect")["risks"]
    p = O.frame_A(f64, "predict", pattern_codes=pat)["risks"]
    det = np.sort(np.concatenate([d, p[p["offset"] < 0]]), order=["i", "j"])
    return det, p[p["offset"] >= 0]


def _check_fused(e, frame, pat):
    from rcd_b200.host import _native as N
    e.upload(frame)
    e.set_patterns(pat)
    e.step(N.MODE_PREDICT, with_detect=True)
    got = e.download()
    det, pred = _fused_oracle(frame, pat)
    compare_pairs(got[got["predicted"] == 0], det, "detect")
    compare_pairs(got[got["predicted"] == 1], pred, "predict")
    c = e.counts()
    assert c["n_pairs"] == len(det) + len(pred) and c["n_fallback"] == 0
    return got


def test_predict_far_from_the_origin_and_outside_static_bounds():
    """fp32 slack of the capsule / window tests at 90 km coordinates (ulp = 8 mm), and predict queries whose
    capsules reach beyond the configured grid (clamped border cells)."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 6000
    frame = W.uniform_frame(n, 81, map_size=1400.0, drone_fraction=0.3)
    for k in ("px", "py"):
        frame[k] = (frame[k] + np.float32(90000.0)).astype(np.float32)
    pat = W.random_patterns(n, 82, p=(0.1, 0.3, 0.5, 0.1))
    for bounds in (None, ((90400.0, 90400.0, 20.0), (91000.0, 91000.0, 60.0))):
        with FrameEngine(n, 1 << 21, world_bounds=bounds) as e:
            got = _check_fused(e, frame, pat)
            assert (got["predicted"] == 1).sum() > 300


def test_fast_and_hard_accelerating_objects():
    """Capsules several hundred metres long, chord slack of tens of metres, windows bent by 3 m/s^2."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 5000
    frame = W.uniform_frame(n, 83, map_size=2500.0, drone_fraction=0.3)
    rng = np.random.default_rng(84)
    for k in ("vx", "vy"):
        frame[k] = (frame[k] * rng.uniform(0.0, 4.0, n)).astype(np.float32)  # up to ~80 m/s
    for k in ("ax", "ay", "az"):
        frame[k] = rng.uniform(-3.0, 3.0, n).astype(np.float32)
    pat = W.random_patterns(n, 85, p=(0.05, 0.25, 0.65, 0.05))
    with FrameEngine(n, 1 << 21) as e:
        got = _check_fused(e, frame, pat)
        assert (got["predicted"] == 1).sum() > 100


def test_queue_overflow_is_finished_in_place():
    """Inter-kernel queues hold 2 * max_pairs + 64 Ki entries; beyond that every stage finishes its pairs
    in place.  Totals stay exact, the first max_pairs records are real records."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 12000
    frame = W.uniform_frame(n, 86, map_size=520.0, drone_fraction=0.3)  # ~1400 neighbours within 100 m each
    pat = W.random_patterns(n, 87, p=(0.1, 0.3, 0.5, 0.1))
    det, pred = _fused_oracle(frame, pat)
    assert len(pred) > 200_000  # far more queue entries than the 65 k + 128 the tiny buffers give
    with FrameEngine(n, 64) as small, FrameEngine(n, 1 << 22) as big:
        for e in (small, big):
            e.upload(frame)
            e.set_patterns(pat)
        big.step(N.MODE_PREDICT, with_detect=True)
        everything = big.download()
        for fused in (True, False):
            if fused:
                small.step(N.MODE_PREDICT, with_detect=True)
            else:
                small.step(N.MODE_DETECT)
                small.step(N.MODE_PREDICT, append=True)
            c = small.counts()
            assert c["n_pairs"] == len(det) + len(pred) == len(everything) and c["n_written"] == 64
            assert c["n_high_risk"] == big.counts()["n_high_risk"] and c["n_alerts"] == big.counts()["n_alerts"]
            got = small.download(sort=False)
            keys = set(zip(everything["i"].tolist(), everything["j"].tolist(), everything["predicted"].tolist()))
            assert len(got) == 64 and all((int(r["i"]), int(r["j"]), int(r["predicted"])) in keys for r in got)


def test_non_finite_state_does_not_disturb_the_rest():
    """The reference raises inside get_grid_id for NaN / inf coordinates (int(nan)) and its caller drops the update;
    here such objects simply never pair with anything, and every other object's result is untouched."""
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 3000
    frame = W.uniform_frame(n, 88, map_size=700.0, drone_fraction=0.3)
    pat = W.random_patterns(n, 89)
    bad = np.arange(0, n, 97)
    clean = np.setdiff1d(np.arange(n), bad)
    dirty = {k: v.copy() for k, v in frame.items()}
    dirty["px"][bad[0::3]] = np.nan
    dirty["vx"][bad[1::3]] = np.inf
    dirty["ay"][bad[2::3]] = np.nan
    with FrameEngine(n, 1 << 21) as e:
        sub = W.take(frame, clean)
        want = _check_fused(e, sub, pat[clean])
        from rcd_b200.host import _native as N
        e.upload(dirty)
        e.set_patterns(pat)
        e.step(N.MODE_PREDICT, with_detect=True)
        got = e.download()
        got = got[~np.isin(got["i"], bad) & ~np.isin(got["j"], bad)]
        remap = np.full(n, -1, np.int64)
        remap[clean] = np.arange(len(clean))
        g = got.copy()
        g["i"], g["j"] = remap[got["i"]], remap[got["j"]]
        assert g.tobytes() == want.tobytes()


def test_pipelined_download_delivers_each_frame_while_the_next_one_runs():
    """rcd_download_begin / _finish: frame k arrives complete although frame k + 1 was uploaded and stepped
    in between (twin pair buffers and totals); the synchronous calls keep seeing the newest frame."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 5000
    frames = [W.uniform_frame(n, 91 + k, map_size=900.0, drone_fraction=0.3) for k in range(4)]
    pat = W.random_patterns(n, 95)
    with FrameEngine(n, 1 << 20) as ref, FrameEngine(n, 1 << 20) as e:
        want = []
        for f in frames:
            ref.upload(f); ref.set_patterns(pat); ref.step(N.MODE_PREDICT, with_detect=True)
            want.append((ref.download(), ref.counts()))
        bufs = [np.zeros(1 << 20, dtype=N.PAIR_DTYPE) for _ in range(2)]
        got = []
        for k, f in enumerate(frames):
            e.upload(f); e.set_patterns(pat); e.step(N.MODE_PREDICT, with_detect=True)
            if k:
                pairs, counts = e.download_finish()  # frame k - 1, after frame k was already enqueued
                got.append((pairs.copy(), counts))   # (the host buffer is reused two frames later)
            e.download_begin(bufs[k % 2])
            with pytest.raises(N.NativeError):
                e.download_begin(bufs[k % 2])        # one delivery at a time
            assert e.counts()["n_pairs"] == want[k][1]["n_pairs"]  # newest frame through the synchronous calls
        got.append(e.download_finish())
        for (pairs, counts), (w_pairs, w_counts) in zip(got, want):
            p = np.sort(pairs, order=["i", "j", "predicted"], kind="stable")
            assert p.tobytes() == w_pairs.tobytes()
            for key in ("n_pairs", "n_candidates", "n_potential", "n_high_risk", "n_alerts", "n_objects"):
                assert counts[key] == w_counts[key], key
        with pytest.raises(N.NativeError):
            e.download_finish()


def test_graph_replay_gives_the_same_frames():
    """RCD_FLAG_GRAPH: after a shape has been seen twice, rcd_step replays a captured CUDA graph.  Results must
    not depend on it -- new data every frame, pipelined delivery (alternating pair buffers), append steps."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    n = 5000
    bounds = ((0.0, 0.0, 0.0), (900.0, 900.0, 100.0))
    frames = [W.uniform_frame(n, 101 + k, map_size=900.0, drone_fraction=0.3) for k in range(8)]
    pat = W.random_patterns(n, 111)
    with FrameEngine(n, 1 << 20, world_bounds=bounds) as ref, FrameEngine(n, 1 << 20, world_bounds=bounds, graph=True) as e:
        bufs = [np.zeros(1 << 20, dtype=N.PAIR_DTYPE) for _ in range(2)]
        got, want = [], []
        for k, f in enumerate(frames):
            ref.upload(f); ref.set_patterns(pat); ref.step(N.MODE_PREDICT, with_detect=True)
            want.append((ref.download(), ref.counts()))
            e.upload(f); e.set_patterns(pat); e.step(N.MODE_PREDICT, with_detect=True)
            if k:
                pairs, counts = e.download_finish()
                got.append((pairs.copy(), counts))
            e.download_begin(bufs[k % 2])
        pairs, counts = e.download_finish()
        got.append((pairs.copy(), counts))
        assert e.graph_replays() >= 3  # two buffer sets: each shape is seen, captured, then replayed
        for (pairs, counts), (w_pairs, w_counts) in zip(got, want):
            assert np.sort(pairs, order=["i", "j", "predicted"], kind="stable").tobytes() == w_pairs.tobytes()
            for key in ("n_pairs", "n_candidates", "n_potential", "n_high_risk", "n_alerts"):
                assert counts[key] == w_counts[key], key
        # two-step frames (detect, then predict appended), other radius, repeated: still identical
        before = e.graph_replays()
        for k in range(4):
            f = frames[k]
            for eng in (ref, e):
                eng.upload(f); eng.set_patterns(pat)
                eng.step(N.MODE_DETECT, 60.0, 4.0)
                eng.step(N.MODE_PREDICT, append=True)
            assert e.download().tobytes() == ref.download().tobytes()
            assert e.counts() == ref.counts()
        assert e.graph_replays() > before


@pytest.mark.parametrize("workload,n", [("cfg1_1k_city", 1000), ("cfg2_5k_city", 5000)])
def test_reference_perf_test_configs_full_oracle(workload, n):
    """BASELINE configs[0] / configs[1] exactly as bench.py runs them (the reference generator's frame and the frame
    after one motion step, performance_test.py:82-195): the fused bench frame against the full oracle, with the
    bench's patterns (all accelerating) and with mixed patterns."""
    import bench
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FrameEngine
    frames, _desc, bounds, _side = bench.make_frames(workload, 1, n, 2)
    with FrameEngine(n, 64 * n, world_bounds=bounds) as e:
        for k, frame in enumerate(frames):
            got = _check_fused(e, frame, np.full(n, 2, np.uint8))
            assert len(got) > 100
            _check_fused(e, frame, W.random_patterns(n, 40 + k))


@pytest.mark.parametrize("law", ["reference", "uniform"])
def test_reduced_configs4_heavy_skew_sampled_oracle_and_overflow(law):
    """configs[4] (10 M objects, Zipf-weighted hotspots, both radial laws of SURVEY.md 8d) reduced to 150 k objects
    at the same density law: a hotspot core with thousands of objects inside one search radius.  (a) sampled queries
    against the full-index oracle, records and per-object risk counts; (b) the same frame through an engine whose
    pair buffer and queues are far too small: the overflow passes finish the pairs in place, every total stays exact
    and the records that were stored are real ones."""
    from rcd_b200.host import _native as N, workloads as W
    from rcd_b200.host.engine import FrameEngine
    O = _oracle()
    n = 150_000
    side = 100000.0 * np.sqrt(n / 10_000_000)
    frame = W.hotspot_frame(n, 2003, side, 3, zipf_s=1.0, radial_law=law)
    pat = np.full(n, 2, np.uint8)
    f64 = f64_frame(frame)
    stride = 101
    bounds = ((0, 0, 0), (side, side, 100))
    with FrameEngine(n, 8_000_000, world_bounds=bounds) as big, FrameEngine(n, 100_000, world_bounds=bounds) as small:
        for e in (big, small):
            e.upload(frame)
            e.set_patterns(pat)
            e.step(N.MODE_PREDICT, with_detect=True)
        cb, cs = big.counts(), small.counts()
        assert cb["n_written"] == cb["n_pairs"], "the big buffer must hold the frame"
        if law == "reference":
            assert cb["n_pairs"] > 1_500_000  # the dense core really is there
        got = big.download()
        sel = got[got["i"] % stride == 0]
        det = O.frame_A(f64, "detect", want_potentials=False, query_stride=stride, risk_cap=1 << 24)["risks"]
        pred = O.frame_A(f64, "predict", pattern_codes=pat, want_potentials=False, query_stride=stride, risk_cap=1 << 24)["risks"]
        compare_pairs(sel[sel["predicted"] == 0], det, "detect")
        compare_pairs(sel[sel["predicted"] == 1], pred, "predict")
        q = np.arange(0, n, stride)
        want_counts = np.bincount(np.concatenate([det["i"], pred["i"]]).astype(np.int64), minlength=n)
        assert np.array_equal(big.risk_counts()[q], want_counts[q])
        # (b) overflow: totals exact, per-object risk counts exact, stored records are a subset
        assert cs["n_written"] == 100_000 < cs["n_pairs"]
        for k in ("n_pairs", "n_candidates", "n_potential", "n_high_risk", "n_alerts"):
            assert cs[k] == cb[k], k
        assert cs["n_fallback"] == 0 and cb["n_fallback"] == 0
        assert np.array_equal(small.risk_counts(), big.risk_counts())
        part = small.download(sort=False)
        key = lambda a: (a["i"].astype(np.int64) << 33) | (a["j"].astype(np.int64) << 1) | a["predicted"].astype(np.int64)
        assert np.isin(key(part), key(got)).all()
