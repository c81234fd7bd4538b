"""Multi-GPU parity test (not collected by pytest: run under torchrun on N GPUs,
`python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py`): slabs + NCCL halo exchange, the
union of the owners' pairs must equal the single-domain oracle result."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from rcd_b200.host import workloads as W, _native as N, slabs as S
from rcd_b200.host.engine import FrameEngine
from tests.gpu_helpers import compare_pairs
from oracle import oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, box = 60000, 5000.0
frame = W.hotspot_frame(n, 77, box, 6, radius_range=(300.0, 600.0))
pat = W.random_patterns(n, 78)
ids = np.arange(n, dtype=np.uint32)
lo, hi = S.slab_bounds(frame, world, box)
halo = S.halo_width([frame])
mine = S.owner_of(frame["px"], lo, hi) == rank
eng = FrameEngine(n, 64 * n, device=local)
stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=torch.device("cuda", local))
ex = S.SlabExchange(eng, lo, hi, rank, world, halo, stream, cap_records=n)
res = {}
for mode in ("detect", "predict"):
    eng.upload(W.take(frame, mine), ids=ids[mine])
    eng.set_patterns(pat[mine])
    nh = ex.exchange()
    got = eng.detect() if mode == "detect" else eng.predict()
    c = eng.counts()
    assert c["n_owned"] == int(mine.sum()) and c["n_objects"] == int(mine.sum()) + nh
    gathered = [None] * world
    dist.all_gather_object(gathered, got)
    if rank == 0:
        both = np.sort(np.concatenate(gathered), order=["i", "j"])
        ora = O.frame_A(W.frame_to_f64(frame), mode, pattern_codes=pat if mode == "predict" else None, threads=8)["risks"]
        compare_pairs(both, ora, mode)
        print(f"{mode}: {len(both)} pairs over {world} slabs == oracle; halo on rank 0: {nh}", flush=True)
# fused frame (detect + predict in one sweep): the union over the owners equals the two oracle results together
eng.upload(W.take(frame, mine), ids=ids[mine])
eng.set_patterns(pat[mine])
ex.exchange()
eng.step(N.MODE_PREDICT, with_detect=True)
got = eng.download()
gathered = [None] * world
dist.all_gather_object(gathered, got)
if rank == 0:
    both = np.concatenate(gathered)
    for mode, flag in (("detect", 0), ("predict", 1)):
        part = np.sort(both[both["predicted"] == flag], order=["i", "j"])
        f64 = W.frame_to_f64(frame)
        ora = O.frame_A(f64, mode, pattern_codes=pat if mode == "predict" else None, threads=8)["risks"]
        if mode == "predict":  # objects without history fall back to detect: those risks are not `predicted`
            ora = ora[ora["offset"] >= 0]
        else:
            nohist = O.frame_A(f64, "predict", pattern_codes=pat, threads=8)["risks"]
            ora = np.sort(np.concatenate([ora, nohist[nohist["offset"] < 0]]), order=["i", "j"])
        compare_pairs(part, ora, mode)
    print(f"fused: {len(both)} pairs over {world} slabs == detect oracle + predict oracle", flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
