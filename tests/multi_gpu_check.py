"""Multi-GPU parity check, run under torchrun on N GPUs (tests/test_gpu_multi.py spawns it with N = 2):
`python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py`.  x-slabs + the NCCL halo exchange
without host round trips (fixed regions, ghost slots): the union of the owners' pairs must equal the single-domain
oracle result -- with the initial cuts and again after the cuts have been moved (re-balancing)."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcd_b200.host import workloads as W, _native as N, slabs as S
from rcd_b200.host.data_sharding import ShardManager
from rcd_b200.host.engine import FrameEngine
from rcd_b200.host.models import Position
from rcd_b200.host.spatial_index import SpatialIndex, SpatialPartitioner
from tests.gpu_helpers import compare_pairs
from oracle import oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, box = 60000, 5000.0
frame = W.hotspot_frame(n, 77, box, 6, radius_range=(300.0, 600.0))
pat = W.random_patterns(n, 78)
ids = np.arange(n, dtype=np.uint32)
halo = S.halo_width([frame])
f64 = W.frame_to_f64(frame)
ora = {}
if rank == 0:
    ora["detect"] = O.frame_A(f64, "detect", threads=8)["risks"]
    ora["predict"] = O.frame_A(f64, "predict", pattern_codes=pat, threads=8)["risks"]


def gather(got):
    parts = [None] * world
    dist.all_gather_object(parts, got)
    return np.concatenate(parts)


def check(lo, hi, label):
    mine = S.owner_of(frame["px"], lo, hi) == rank
    eng = FrameEngine(n + 1024 * world, 64 * n, device=local)
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=torch.device("cuda", local))
    ex = S.SlabExchange(eng, lo, hi, rank, world, halo, stream)
    eng.upload(W.take(frame, mine), ids=ids[mine])
    eng.set_patterns(pat[mine])
    ex.configure()
    for mode in ("detect", "predict"):
        eng.upload(W.take(frame, mine), ids=ids[mine])
        eng.set_patterns(pat[mine])
        nh = ex.exchange()
        got = eng.detect() if mode == "detect" else eng.predict()
        assert not ex.overflowed(), "halo region too small"
        c = eng.counts()
        assert c["n_owned"] == int(mine.sum()) and c["n_objects"] == int(mine.sum()) + nh
        both = gather(got)
        if rank == 0:
            compare_pairs(np.sort(both, order=["i", "j"]), ora[mode], mode)
            print(f"{label} {mode}: {len(both)} pairs over {world} slabs == oracle; halo slots on rank 0: {nh} "
                  f"(sent {ex.sent_counts().tolist()})", flush=True)
    # fused frame (detect + predict in one sweep): the union over the owners equals the two oracle results together
    eng.upload(W.take(frame, mine), ids=ids[mine])
    eng.set_patterns(pat[mine])
    ex.exchange()
    eng.step(N.MODE_PREDICT, with_detect=True)
    both = gather(eng.download())
    if rank == 0:
        for mode, flag in (("detect", 0), ("predict", 1)):
            part = np.sort(both[both["predicted"] == flag], order=["i", "j"])
            if mode == "predict":  # objects without history fall back to detect: those risks are not `predicted`
                want = ora["predict"][ora["predict"]["offset"] >= 0]
            else:
                nohist = ora["predict"][ora["predict"]["offset"] < 0]
                want = np.sort(np.concatenate([ora["detect"], nohist]), order=["i", "j"])
            compare_pairs(part, want, mode)
        print(f"{label} fused: {len(both)} pairs over {world} slabs == detect oracle + predict oracle", flush=True)
    # a region that is too small drops records and says so
    small = S.SlabExchange(eng, lo, hi, rank, world, halo, stream, slack=1.0, min_records=0)
    eng.upload(W.take(frame, mine), ids=ids[mine])
    small.configure()
    small.cap_send = [max(0, c // 2) for c in small.cap_send]  # (both sides shrink alike: cap_recv mirrors the peers' cap_send)
    small.cap_recv = [max(0, c // 2) for c in small.cap_recv]
    small.send_offset = np.concatenate([[0], np.cumsum(small.cap_send)]).astype(np.uint64)
    small.n_recv = int(sum(small.cap_recv))
    eng.upload(W.take(frame, mine), ids=ids[mine])
    small.exchange()
    eng.sync()
    assert small.overflowed() == bool(sum(small.sent_counts()) > 0), "overflow of a halo region went unnoticed"
    eng.close()
    dist.barrier()


lo, hi = S.slab_bounds(frame, world, box)
check(lo, hi, "initial cuts")
# The shard-manager adapter drives the re-balancing: the partitioner holds the cuts (slab mode),
# the manager routes every vehicle to the slab it lies in and feeds the slabs' loads back; rebalance_shards() moves
# the cuts (pretend the first slab was the slowest by far) and the halo exchange follows them.
index = SpatialIndex()
for k in range(n):
    index.insert_vehicle(f"v{k}", Position(float(frame["px"][k]), float(frame["py"][k]), float(frame["pz"][k])))
part = SpatialPartitioner(index, num_shards=world)


class _Follower:  # stands for this rank's SlabExchange: receives the cuts the partitioner decides on
    cuts = None

    def set_cuts(self, lo, hi):
        self.cuts = (lo, hi)


follow = _Follower()
part.attach_slabs(lo, hi, box, exchanges=[follow])
mgr = ShardManager(part, initial_shards=world)
routed = np.array([int(mgr.get_shard_for_vehicle(f"v{k}", index.get_vehicle_position(f"v{k}"))[6:]) for k in range(0, n, 7)])
assert np.array_equal(routed, S.owner_of(frame["px"][::7], lo, hi)), "ShardManager routes by slab"
mgr.update_shard_loads([3.0] + [1.0] * (world - 1))
res = part.rebalance_shards()
assert res["cuts_moved"] and follow.cuts is not None
lo2, hi2 = follow.cuts
assert not np.array_equal(hi, hi2) and float(hi2[0]) < float(hi[0]), "the slow slab must shrink"
moved = sum(mgr.get_shard_for_vehicle(f"v{k}", index.get_vehicle_position(f"v{k}")) != f"shard-{routed[i]}"
            for i, k in enumerate(range(0, n, 7)))
assert moved == mgr.migrations > 0, "vehicles between the old and the new cut migrate"
check(lo2, hi2, "re-balanced cuts")
if rank == 0:
    print(f"shard manager: {moved} of {len(routed)} sampled vehicles migrated after rebalance_shards(); "
          f"cuts {[round(float(v), 1) for v in hi[:-1]]} -> {[round(float(v), 1) for v in hi2[:-1]]}", flush=True)
    print("MULTI_GPU_CHECK_OK", flush=True)
dist.destroy_process_group()
