"""Spatial slabs over the GPUs of one box.

The reference shards *queries* over nodes while every node indexes every vehicle
(src/collision_system.py:437-446, src/collision/data_sharding.py:172-201).  Here each GPU owns
the objects of one x-slab and receives, every frame, copies of the objects of the other slabs
that lie within the halo width of its slab; results are emitted by the owner of the querying
object only, so the union over GPUs equals the single-domain result and no reduction is needed.

Host-side pieces (numpy, used by the CPU/gloo tests and by bench.py) and the device exchange
(packed on the GPU by ``rcd_halo_pack``, moved with one NCCL all_to_all, appended by
``rcd_halo_append``).  Halo record = 13 x 32-bit words: the 11 fp32 state fields, meta
(type | pattern << 8) and the caller id.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FRAME_FIELDS = ("px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az", "size", "heading")
RECORD_WORDS = 13


def slab_bounds(frame: Dict[str, np.ndarray], n_slabs: int, side: float, cell: float = 100.0,
                pair_weight: float = 0.25) -> Tuple[np.ndarray, np.ndarray]:
    """x-slabs of equal estimated work: every object weighs 1 + pair_weight * (population of its
    cell), since pair work grows with the local density."""
    if n_slabs == 1:
        return np.array([-np.inf], np.float32), np.array([np.inf], np.float32)
    x, y = frame["px"].astype(np.float64), frame["py"].astype(np.float64)
    nc = max(1, int(side / cell))
    cx = np.clip((x / cell).astype(np.int64), 0, nc - 1)
    cy = np.clip((y / cell).astype(np.int64), 0, nc - 1)
    pop = np.bincount(cx * nc + cy, minlength=nc * nc)
    w = 1.0 + pair_weight * pop[cx * nc + cy]
    order = np.argsort(x, kind="stable")
    cw = np.cumsum(w[order])
    cuts = []
    for k in range(1, n_slabs):
        at = min(len(order) - 1, int(np.searchsorted(cw, cw[-1] * k / n_slabs)))
        cuts.append(float(np.float32(x[order[at]])))
    cuts = sorted(cuts)
    lo = np.array([-np.inf] + cuts, np.float32)
    hi = np.array(cuts + [np.inf], np.float32)
    return lo, hi


def halo_width(frames: Sequence[Dict[str, np.ndarray]], radius: float = 100.0, predict: bool = True) -> float:
    """Search radius plus the farthest a predicted centre can move in 9.5 s (quirk Q5: predicted
    self positions are compared with the others' current positions)."""
    h = radius
    if predict:
        for f in frames:
            if len(f["px"]) == 0:
                continue
            v = np.sqrt(f["vx"].astype(np.float64) ** 2 + f["vy"].astype(np.float64) ** 2 + f["vz"].astype(np.float64) ** 2).max()
            a = np.sqrt(f["ax"].astype(np.float64) ** 2 + f["ay"].astype(np.float64) ** 2 + f["az"].astype(np.float64) ** 2).max()
            h = max(h, 100.0 + 9.5 * v + 45.125 * a)
    return float(h * 1.001 + 0.5)


def owner_of(x: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    """Slab index of every x (slabs are half-open [lo, hi))."""
    return np.clip(np.searchsorted(np.asarray(hi, np.float32), np.asarray(x, np.float32), side="right"), 0, len(lo) - 1)


def pack_halo_numpy(frame: Dict[str, np.ndarray], ids: np.ndarray, pattern: Optional[np.ndarray], lo, hi,
                    rank: int, halo: float) -> Tuple[np.ndarray, np.ndarray]:
    """Host mirror of rcd_halo_pack: records of the owned objects every peer needs, grouped by peer."""
    n = len(frame["px"])
    x = frame["px"].astype(np.float32)
    pat = np.full(n, 2, np.uint8) if pattern is None else np.asarray(pattern, np.uint8)
    recs, counts = [], np.zeros(len(lo), np.int64)
    h = np.float32(halo)
    for p in range(len(lo)):
        if p == rank:
            continue
        m = (x >= np.float32(lo[p]) - h) & (x < np.float32(hi[p]) + h)
        k = int(m.sum())
        counts[p] = k
        if k == 0:
            continue
        r = np.zeros((k, RECORD_WORDS), np.uint32)
        for c, name in enumerate(FRAME_FIELDS):
            r[:, c] = frame[name][m].astype(np.float32).view(np.uint32)
        r[:, 11] = frame["type"][m].astype(np.uint32) | (pat[m].astype(np.uint32) << 8)
        r[:, 12] = np.asarray(ids, np.uint32)[m]
        recs.append(r)
    rec = np.concatenate(recs) if recs else np.zeros((0, RECORD_WORDS), np.uint32)
    return rec, counts


def unpack_halo_numpy(rec: np.ndarray):
    """records -> (frame, ids, pattern)."""
    rec = np.asarray(rec, np.uint32).reshape(-1, RECORD_WORDS)
    frame = {name: np.ascontiguousarray(rec[:, c]).view(np.float32) for c, name in enumerate(FRAME_FIELDS)}
    frame["type"] = (rec[:, 11] & 0xFF).astype(np.uint8)
    pattern = ((rec[:, 11] >> 8) & 0xFF).astype(np.uint8)
    return frame, rec[:, 12].copy(), pattern


def all_to_all_records(send, send_counts: Sequence[int], group=None):
    """Variable-size all_to_all of halo records (torch tensor [m, 13] int32, grouped by peer).
    NCCL: one all_to_all_single; gloo (CPU tests): pairwise isend / irecv."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sc = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=send.device)
    rc = torch.empty_like(sc)
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(rc, sc, group=group)
    else:
        gathered = [torch.empty_like(sc) for _ in range(world)]
        dist.all_gather(gathered, sc, group=group)
        rc = torch.stack([g[rank] for g in gathered])
    recv_counts = [int(v) for v in rc.cpu()]
    recv = torch.empty((sum(recv_counts), send.shape[1]), dtype=send.dtype, device=send.device)
    if backend == "nccl":
        dist.all_to_all_single(recv, send[: int(sum(send_counts))], output_split_sizes=recv_counts,
                               input_split_sizes=[int(c) for c in send_counts], group=group)
    else:
        ops, so, ro = [], 0, 0
        for p in range(world):
            if p != rank and send_counts[p]:
                ops.append(dist.P2POp(dist.isend, send[so: so + int(send_counts[p])].contiguous(), p, group=group))
            if p != rank and recv_counts[p]:
                ops.append(dist.P2POp(dist.irecv, recv[ro: ro + recv_counts[p]], p, group=group))
            so += int(send_counts[p])
            ro += recv_counts[p]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    return recv, recv_counts


def rebalanced_cuts(x: np.ndarray, lo: np.ndarray, hi: np.ndarray, ms: Sequence[float], side: float,
                    damping: float = 0.7) -> Tuple[np.ndarray, np.ndarray]:
    """New x-cuts from the measured frame time of every slab (the re-balancing step of SURVEY.md 8e): every object
    of slab r weighs ms[r] / count[r] (its slab's measured cost per object), the cuts move to equal cumulative
    weight.  `damping` < 1 moves only part of the way (the cost per object changes with the cut)."""
    n_slabs = len(lo)
    if n_slabs == 1:
        return np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    x = np.asarray(x, np.float64)
    owner = owner_of(x, lo, hi)
    count = np.bincount(owner, minlength=n_slabs).astype(np.float64)
    per_obj = np.asarray(ms, np.float64) / np.maximum(count, 1.0)
    w = per_obj[owner]
    order = np.argsort(x, kind="stable")
    cw = np.cumsum(w[order])
    old = np.asarray(hi[:-1], np.float64)
    cuts = []
    for k in range(1, n_slabs):
        at = min(len(order) - 1, int(np.searchsorted(cw, cw[-1] * k / n_slabs)))
        cuts.append(float(x[order[at]]))
    cuts = np.asarray(cuts, np.float64)
    cuts = old + damping * (cuts - old)
    cuts = np.maximum.accumulate(cuts)
    cuts = [float(np.float32(c)) for c in cuts]
    return np.array([-np.inf] + cuts, np.float32), np.array(cuts + [np.inf], np.float32)


class SlabExchange:
    """Per-frame halo exchange of one rank on the GPU, without a host round trip: single-pass pack into fixed
    per-peer regions (``rcd_halo_pack_async``; unused slots stay ghost records) -> fixed-size NCCL all_to_all on
    the engine's stream -> append of everything received (``rcd_halo_append``).  The region sizes come from a
    counting pass over a representative frame (``configure``) and carry slack; ``overflowed()`` (which does wait
    for the device) tells whether a region was too small and the sizes have to be taken again."""

    def __init__(self, engine, lo, hi, rank: int, world: int, halo: float, stream, group=None, slack: float = 1.5,
                 min_records: int = 256):
        self.engine, self.rank, self.world, self.halo = engine, int(rank), int(world), float(halo)
        self.lo, self.hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
        self.stream, self.group = stream, group
        self.slack, self.min_records = float(slack), int(min_records)
        self.launches_last = 0
        self.halo_last = 0
        self.send = self.recv = self.counts = None

    def set_cuts(self, lo, hi) -> None:
        self.lo, self.hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)

    def configure(self) -> None:
        """Size the per-peer regions from the objects the engine holds now (its owned objects of a representative
        frame): a counting pass on this rank, one exchange of the sizes with the peers.  Waits for the device."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.engine.device)
        counts = self.engine.halo_pack(self.lo, self.hi, self.rank, self.halo, 0, 0)  # counting pass only
        cap_send = [0 if p == self.rank else int(c * self.slack) + self.min_records for p, c in enumerate(counts)]
        sc = torch.tensor(cap_send, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        self.cap_send, self.cap_recv = cap_send, [int(v) for v in rc.cpu()]
        self.send_offset = np.concatenate([[0], np.cumsum(self.cap_send)]).astype(np.uint64)
        self.n_recv = int(sum(self.cap_recv))
        self.send = torch.empty((max(1, int(self.send_offset[-1])), RECORD_WORDS), dtype=torch.int32, device=dev)
        self.recv = torch.empty((max(1, self.n_recv), RECORD_WORDS), dtype=torch.int32, device=dev)
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self.halo_last = self.n_recv

    def exchange(self) -> int:
        """Pack this rank's boundary objects, trade them, append what the peers sent.  The engine must hold only
        its owned objects (call after upload).  Returns the halo slots appended (ghosts included)."""
        import torch
        import torch.distributed as dist
        self.engine.halo_pack_async(self.lo, self.hi, self.rank, self.halo, self.send.data_ptr(), self.send_offset,
                                    self.counts.data_ptr())
        with torch.cuda.stream(self.stream):
            dist.all_to_all_single(self.recv[: self.n_recv], self.send[: int(self.send_offset[-1])],
                                   output_split_sizes=self.cap_recv, input_split_sizes=self.cap_send, group=self.group)
        self.engine.halo_append(self.recv.data_ptr() if self.n_recv else 0, self.n_recv)
        self.launches_last = 1 + (1 if self.n_recv else 0)
        return self.n_recv

    def sent_counts(self) -> np.ndarray:
        """Records this rank wanted to send to every peer in the last exchange (waits for the device)."""
        return self.counts.cpu().numpy().astype(np.int64)

    def overflowed(self) -> bool:
        return bool(np.any(self.sent_counts() > np.asarray(self.cap_send, np.int64)))
