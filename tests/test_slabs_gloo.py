"""Multi-rank slab logic on CPU: world_size-2 (and 3) gloo process groups exchange halo records;
the oracle stands in for the GPU frame.  The union of the owners' results must equal the
single-domain result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, box, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from rcd_b200.host import slabs as S
        from rcd_b200.host import workloads as W
        frame = W.uniform_frame(n, 5, map_size=box, drone_fraction=0.3)
        pattern = W.random_patterns(n, 6, p=(0.1, 0.4, 0.5, 0.0))
        ids = np.arange(n, dtype=np.uint32)
        lo, hi = S.slab_bounds(frame, world, box)
        halo = S.halo_width([frame])
        mine = S.owner_of(frame["px"], lo, hi) == rank
        own, own_ids, own_pat = W.take(frame, mine), ids[mine], pattern[mine]
        rec, counts = S.pack_halo_numpy(own, own_ids, own_pat, lo, hi, rank, halo)
        recv, recv_counts = S.all_to_all_records(torch.from_numpy(rec.view(np.int32)), counts)
        assert recv_counts[rank] == 0
        hf, hids, hpat = S.unpack_halo_numpy(recv.numpy().view(np.uint32))
        full = W.concat(own, hf)
        full_ids = np.concatenate([own_ids, hids])
        full_pat = np.concatenate([own_pat, hpat])
        n_own = len(own_ids)
        out = {}
        for mode in ("detect", "predict"):
            r = O.frame_A(W.frame_to_f64(full), mode, pattern_codes=full_pat if mode == "predict" else None)["risks"]
            r = r[r["i"] < n_own]  # results are emitted by the owner of the querying object only
            out[mode] = np.stack([full_ids[r["i"]], full_ids[r["j"]], r["ttc"], r["risk"]], 1)
        gathered = [None] * world
        dist.all_gather_object(gathered, (out, n_own, len(hids)))
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_with_halo_equal_single_domain(world):
    from oracle import oracle as O
    from rcd_b200.host import workloads as W
    n, box = 3000, 2500.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, box, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frame = W.uniform_frame(n, 5, map_size=box, drone_fraction=0.3)
    pattern = W.random_patterns(n, 6, p=(0.1, 0.4, 0.5, 0.0))
    assert sum(g[1] for g in gathered) == n          # every object has exactly one owner
    assert all(g[2] > 0 for g in gathered)           # and every slab received a halo
    assert sum(g[2] for g in gathered) < n           # ... that is smaller than the frame
    for mode in ("detect", "predict"):
        want = O.frame_A(W.frame_to_f64(frame), mode, pattern_codes=pattern if mode == "predict" else None)["risks"]
        want = np.stack([want["i"], want["j"], want["ttc"], want["risk"]], 1)
        got = np.concatenate([g[0][mode] for g in gathered])
        got = got[np.lexsort((got[:, 1], got[:, 0]))]
        assert len(want) > 10
        assert np.array_equal(got, want), mode


def test_slab_bounds_balance_work_not_counts():
    from rcd_b200.host import slabs as S
    from rcd_b200.host import workloads as W
    frame = W.hotspot_frame(40000, 3, 8000.0, 3, drone_fraction=0.0)
    lo, hi = S.slab_bounds(frame, 4, 8000.0)
    assert np.all(lo[1:] == hi[:-1]) and lo[0] == -np.inf and hi[-1] == np.inf
    assert np.all(np.diff(hi[:-1]) >= 0)
    owner = S.owner_of(frame["px"], lo, hi)
    assert set(owner.tolist()) == {0, 1, 2, 3}
    # every object lands in the slab whose half-open interval contains it
    assert np.all((frame["px"] >= lo[owner]) & (frame["px"] < hi[owner]))


def test_pack_unpack_roundtrip():
    from rcd_b200.host import slabs as S
    from rcd_b200.host import workloads as W
    frame = W.uniform_frame(500, 9, map_size=1000.0, drone_fraction=0.5)
    ids = np.arange(500, dtype=np.uint32) + 7
    pat = W.random_patterns(500, 1)
    lo, hi = np.array([-np.inf, 500.0], np.float32), np.array([500.0, np.inf], np.float32)
    rec, counts = S.pack_halo_numpy(frame, ids, pat, lo, hi, 0, 120.0)
    m = frame["px"] >= np.float32(500.0 - 120.0)
    assert counts.tolist() == [0, int(m.sum())]
    f2, ids2, pat2 = S.unpack_halo_numpy(rec)
    for k in S.FRAME_FIELDS:
        assert np.array_equal(f2[k], frame[k][m])
    assert np.array_equal(f2["type"], frame["type"][m]) and np.array_equal(ids2, ids[m]) and np.array_equal(pat2, pat[m])


def test_rebalanced_cuts_move_towards_equal_cost():
    """The re-balancing step: with a cost model in which an object's cost grows with the local density,
    repeating `rebalanced_cuts` on the measured slab costs brings the slabs' costs together; the cuts stay ordered and
    every object keeps exactly one owner."""
    from rcd_b200.host import slabs as S
    from rcd_b200.host import workloads as W
    side, n_slabs = 8000.0, 4
    frame = W.hotspot_frame(60000, 5, side, 4, drone_fraction=0.0)
    x = frame["px"].astype(np.float64)
    cell = np.clip((x / 100.0).astype(np.int64), 0, 79) * 80 + np.clip((frame["py"] / 100.0).astype(np.int64), 0, 79)
    cost_obj = 1.0 + 0.5 * np.bincount(cell, minlength=6400)[cell]  # "pair work" of every object

    def slab_cost(lo, hi):
        return np.bincount(S.owner_of(frame["px"], lo, hi), weights=cost_obj, minlength=n_slabs)

    lo = np.array([-np.inf] + [side * k / n_slabs for k in range(1, n_slabs)], np.float32)  # naive equal-width cuts
    hi = np.array([side * k / n_slabs for k in range(1, n_slabs)] + [np.inf], np.float32)
    first = slab_cost(lo, hi)
    for _ in range(6):
        lo, hi = S.rebalanced_cuts(frame["px"], lo, hi, slab_cost(lo, hi), side)
        assert lo[0] == -np.inf and hi[-1] == np.inf and np.all(lo[1:] == hi[:-1]) and np.all(np.diff(hi[:-1]) >= 0)
        owner = S.owner_of(frame["px"], lo, hi)
        assert np.all((frame["px"] >= lo[owner]) & (frame["px"] < hi[owner]))
    last = slab_cost(lo, hi)
    assert last.max() / last.mean() < 1.1 < first.max() / first.mean(), (first, last)
    # one slab: nothing to cut
    l1, h1 = S.rebalanced_cuts(frame["px"], np.array([-np.inf], np.float32), np.array([np.inf], np.float32), [1.0], side)
    assert l1.tolist() == [-np.inf] and h1.tolist() == [np.inf]
