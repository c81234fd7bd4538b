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


class SlabExchange:
    """Per-frame halo exchange of one rank on the GPU: pack (CUDA) -> all_to_all (NCCL) -> append."""

    def __init__(self, engine, lo, hi, rank: int, world: int, halo: float, stream, cap_records: int, group=None):
        import torch
        self.engine, self.rank, self.world, self.halo = engine, int(rank), int(world), float(halo)
        self.lo, self.hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
        self.stream, self.group, self.cap = stream, group, int(cap_records)
        dev = torch.device("cuda", engine.device)
        self.send = torch.empty((self.cap, RECORD_WORDS), dtype=torch.int32, device=dev)
        self.launches_last = 0
        self.halo_last = 0

    def exchange(self) -> int:
        """Pack this rank's boundary objects, trade them, append what the peers sent.  The engine
        must hold only its owned objects (call after upload).  Returns the halo object count."""
        import torch
        counts = self.engine.halo_pack(self.lo, self.hi, self.rank, self.halo, self.send.data_ptr(), self.cap)
        with torch.cuda.stream(self.stream):
            recv, recv_counts = all_to_all_records(self.send, counts, self.group)
        total = int(sum(recv_counts))
        self._keep = recv  # alive until the append kernel has consumed it
        self.engine.halo_append(recv.data_ptr() if total else 0, total)
        self.launches_last = (2 if int(np.sum(counts)) else 1) + (1 if total else 0)
        self.halo_last = total
        return total
