#!/usr/bin/env python
"""Headline benchmark: object-updates/s and p50/p99 frame latency of the collision hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A *step* is one frame = what the reference's perf harness times
(src/test/performance_test.py:794-813): ingest every object, rebuild the spatial index, run
detect_collisions for every object and predict_collisions for every object, deliver the results.

Workload (default): BASELINE.json configs[3] -- 1M mixed vehicles+drones, 3-D, clustered in 50
urban hotspots on a 31.6 km map (reference generator law r = U*radius).  With N GPUs the frame is
N x 1M objects on a map of N times the area (same density, 50 hotspots per 1M objects): weak
scaling.  Space is cut into N x-slabs of equal estimated pair work; every frame each GPU packs the
objects within the halo width of the other slabs and exchanges them with NCCL (all_to_all over
NVLink), then runs the frame on owned + halo objects (results are emitted by the owner only).

value : whole-job object-updates/s with the frame's object state already in HBM when the timed
        region starts (the engine still ingests it device-to-device every frame).  Timed on the
        device with CUDA events on the engine's stream, max over ranks; L2 is flushed (256 MiB
        write) before every timed frame.
e2e   : same metric through the public host API: pinned host SoA arrays -> H2D -> frame -> D2H of
        the totals and the emitted pairs, wall clock, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "object-updates/s"
PER_GPU_DEFAULT = 1_000_000
H2D_BYTES_PER_OBJECT = 11 * 4 + 1 + 1  # 11 fp32 fields + type + trajectory pattern


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_frames(workload: str, n_gpus: int, per_gpu: int, n_frames: int = 2):
    """Global frames (the same on every rank: deterministic seeds) + description."""
    from rcd_b200.host import workloads as W
    n = per_gpu * n_gpus
    if workload == "cfg4_1m_clustered3d":
        side = 31623.0 * np.sqrt(n / 1_000_000)
        f0 = W.hotspot_frame(n, 2002, side, max(1, round(50 * n / 1_000_000)))
        desc = (f"configs[3]: {n} vehicles+drones 3-D, {max(1, round(50 * n / 1_000_000))} hotspots (r=U*radius), "
                f"{side / 1000:.1f} km map")
        bounds = ((0.0, 0.0, 0.0), (side, side, 100.0))
    elif workload in ("cfg5_10m_skew3d", "cfg5_10m_skew3d_uniform_disc"):
        # the reference's radial law r = U * radius piles a Zipf-weighted hotspot up at its centre (density ~ 1/r):
        # 1.8e9 emitted pairs per frame at 10 M objects, more than any pair buffer holds.  SURVEY.md 8d asks for that
        # frame to be reported as it is, next to the variant with objects uniform in each disc (r = sqrt(U) * radius).
        law = "uniform" if workload.endswith("uniform_disc") else "reference"
        side = 100000.0 * np.sqrt(n / 10_000_000)
        f0 = W.hotspot_frame(n, 2003, side, max(1, round(200 * n / 10_000_000)), zipf_s=1.0, radial_law=law)
        desc = (f"configs[4] family: {n} objects 3-D, Zipf(s=1) hotspots ({'r=sqrt(U)*radius' if law == 'uniform' else 'r=U*radius'}), "
                f"{side / 1000:.1f} km map")
        bounds = ((0.0, 0.0, 0.0), (side, side, 100.0))
    elif workload == "cfg3_100k_uniform2d":
        side = 10000.0 * np.sqrt(n / 100_000)
        f0 = W.uniform_frame(n, 2001, map_size=side)
        desc = f"configs[2]: {n} uniform vehicles 2-D, {side / 1000:.1f} km map"
        bounds = ((0.0, 0.0, 0.0), (side, side, 0.0))
    elif workload == "cfg2_5k_city":
        f0 = W.reference_city_frame(n, 1235)
        desc = f"configs[1]: reference perf-test generator, {n} vehicles, 10 km map"
        side = 10000.0
        bounds = ((0.0, 0.0, 0.0), (side, side, 0.0))
    else:
        raise SystemExit(f"unknown workload {workload}")
    frames = [f0]
    rng = np.random.default_rng(99)
    for _ in range(1, n_frames):  # the reference advances every vehicle between frames (:147-195)
        frames.append(W.advance(frames[-1], 0.05, rng, map_size=(side, side)))
    return frames, desc, bounds, side


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference algorithm, all host threads, bounded sample)
# ----------------------------------------------------------------------------------------------
def cpu_frame_seconds(frame, pattern, budget_s: float):
    """Time of one reference frame (index + detect-all + predict-all) on the host cores,
    extrapolated from a bounded sample: every `stride`-th object is queried against the FULL
    index; the index build is timed in full.  Returns (seconds_per_frame, description, threads)."""
    from oracle import oracle as O
    from rcd_b200.host import workloads as W
    f64 = W.frame_to_f64(frame)
    n = len(frame["px"])
    threads = max(O.max_threads(), os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: ask for all cores
    # index-only cost (stride so large that only object 0 is queried), once per mode
    t0 = time.perf_counter()
    O.frame_A(f64, "detect", want_potentials=False, query_stride=max(n, 1), risk_cap=1 << 16, threads=threads)
    t_index = time.perf_counter() - t0
    # pilot to size the sample
    stride = max(1, n // 2000)
    t0 = time.perf_counter()
    O.frame_A(f64, "detect", want_potentials=False, query_stride=stride, risk_cap=1 << 22, threads=threads)
    O.frame_A(f64, "predict", pattern_codes=pattern, want_potentials=False, query_stride=stride, risk_cap=1 << 22, threads=threads)
    t_pilot = max(time.perf_counter() - t0 - 2 * t_index, 1e-6)
    per_query = t_pilot / max(1, (n + stride - 1) // stride)
    want = int(min(n, max(2000, budget_s / per_query)))
    stride = max(1, n // want)
    t0 = time.perf_counter()
    O.frame_A(f64, "detect", want_potentials=False, query_stride=stride, risk_cap=1 << 24, threads=threads)
    O.frame_A(f64, "predict", pattern_codes=pattern, want_potentials=False, query_stride=stride, risk_cap=1 << 24, threads=threads)
    t_sample = time.perf_counter() - t0
    queried = (n + stride - 1) // stride
    t_queries = max(t_sample - 2 * t_index, 1e-9)
    t_frame = 2 * t_index + t_queries * (n / queried)
    desc = (f"oracle/oracle.c (float64 port of src/collision, OpenMP {threads} threads): index built over all "
            f"{n} objects, every {stride}-th object queried (detect + predict = {queried} queries each, "
            f"{t_sample:.1f} s measured), extrapolated to the full frame")
    return t_frame, desc, threads


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference
    itself is pure Python and absent from the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames, desc, _bounds, _side = make_frames(args.workload, args.gpus, args.objects_per_gpu, 1)
    n = len(frames[0]["px"])
    pattern = np.full(n, 2, np.uint8)
    steps = max(1, args.steps)
    budget = max(2.0, min(20.0, 120.0 / (steps + args.warmup)))
    times = []
    info = None
    for k in range(args.warmup + steps):
        t, info, threads = cpu_frame_seconds(frames[0], pattern, budget)
        if k >= args.warmup:
            times.append(t)
    t_mean = float(np.mean(times))
    value = n / t_mean
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "object-updates/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": t_mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "objects": n, "frame": "index + detect-all + predict-all"},
        "cpu_baseline": {"value": value, "unit": "object-updates/s", "cores": threads, "kind": "port", "sample": info},
        "e2e": {"value": value, "unit": "object-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from rcd_b200.host import _native as N
    from rcd_b200.host import workloads as W
    from rcd_b200.host.engine import FRAME_FIELDS, FrameEngine
    from rcd_b200.host.slabs import SlabExchange, halo_width, slab_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    frames, desc, bounds, side = make_frames(args.workload, world, args.objects_per_gpu, 2)
    n_total = len(frames[0]["px"])
    lo, hi = slab_bounds(frames[0], world, side)
    halo = halo_width(frames)
    ids_all = np.arange(n_total, dtype=np.uint32)
    own, own_ids = [], []
    for f in frames:
        m = (f["px"] >= lo[rank]) & (f["px"] < hi[rank])
        own.append(W.take(f, m))
        own_ids.append(ids_all[m])
    n_own_max = max(len(o["px"]) for o in own)
    # exact halo sizes (narrow slabs in dense hotspots send the same object to several peers)
    n_send_max, n_recv_max = 0, 0
    if world > 1:
        for f in frames:
            x = f["px"]
            owner = (np.searchsorted(hi, x, side="right")).clip(0, world - 1)
            mine = owner == rank
            xs = x[mine]
            n_send = sum(int(((xs >= lo[p] - halo) & (xs < hi[p] + halo)).sum()) for p in range(world) if p != rank)
            n_recv = int(((x >= lo[rank] - halo) & (x < hi[rank] + halo) & ~mine).sum())
            n_send_max, n_recv_max = max(n_send_max, n_send), max(n_recv_max, n_recv)
    cap = int(n_own_max + 1.1 * n_recv_max) + 4096
    max_pairs = int(args.max_pairs)
    # slab bounding box (+ halo) as the static grid bounds: no per-frame bbox round trip
    xlo = max(0.0, float(lo[rank]) - halo) if np.isfinite(lo[rank]) else 0.0
    xhi = min(side, float(hi[rank]) + halo) if np.isfinite(hi[rank]) else side
    eng_bounds = ((xlo, bounds[0][1], bounds[0][2]), (xhi, bounds[1][1], bounds[1][2]))
    # --graph: rcd_step is replayed as a CUDA graph (RCD_FLAG_GRAPH), which excludes the per-stage events; the stage
    # breakdown then comes from a few extra frames on a profiled twin engine after the timed regions
    eng = FrameEngine(cap, max_pairs, device=local_rank, world_bounds=eng_bounds, profile=not args.graph, graph=args.graph)
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=torch.device("cuda", local_rank))
    exch = SlabExchange(eng, lo, hi, rank, world, halo, stream, cap_records=int(1.1 * n_send_max) + 4096) if world > 1 else None

    # device-resident copies of the frames (the engine ingests them device-to-device every step)
    dev = []
    pin = []
    for f, fid in zip(own, own_ids):
        n = len(f["px"])
        d = {k: torch.from_numpy(f[k]).cuda() for k in FRAME_FIELDS}
        d["type"] = torch.from_numpy(f["type"]).cuda()
        d["id"] = torch.from_numpy(fid.astype(np.int32)).cuda()
        d["pattern"] = torch.full((n,), 2, dtype=torch.uint8, device="cuda")
        dev.append(d)
        p = {k: torch.from_numpy(f[k]).pin_memory() for k in FRAME_FIELDS}
        p["type"] = torch.from_numpy(f["type"]).pin_memory()
        p["id"] = torch.from_numpy(fid.astype(np.int32)).pin_memory()
        p["pattern"] = torch.full((n,), 2, dtype=torch.uint8).pin_memory()
        pin.append(p)
    pairs_pin = torch.empty(min(max_pairs, int(args.host_pairs_cap)) * 48, dtype=torch.uint8).pin_memory()
    pairs_host = pairs_pin.numpy().view(N.PAIR_DTYPE)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    launches = [0]
    stage_acc = {}

    def frame_resident(k: int):
        d = dev[k % len(dev)]
        n = int(d["px"].shape[0])
        eng.upload_device(n, [d[f].data_ptr() for f in FRAME_FIELDS], d["type"].data_ptr(), d["id"].data_ptr())
        eng.set_patterns_device(n, d["pattern"].data_ptr())
        if exch is not None:
            exch.exchange()
        eng.step(N.MODE_PREDICT, with_detect=True)  # detect-all + predict-all, one sweep

    def frame_e2e(k: int):
        p = pin[k % len(pin)]
        n = int(p["px"].shape[0])
        eng.upload_host_ptrs(n, [p[f].data_ptr() for f in FRAME_FIELDS], p["type"].data_ptr(), p["id"].data_ptr())
        eng.set_patterns_host_ptr(n, p["pattern"].data_ptr())
        if exch is not None:
            exch.exchange()
        eng.step(N.MODE_PREDICT, with_detect=True)  # detect-all + predict-all, one sweep
        return eng.download(sort=False, out=pairs_host)  # delivery only: consumers group on their own

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # runs through warm-up and both timed regions (the frames are milliseconds long)
        time.sleep(0.3)
    for k in range(args.warmup):
        frame_resident(k)
        eng.sync()
    counts = eng.counts()

    # ---- timed region: device-resident frames ---------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush: 256 MiB write, outside the timed events
            ev[k][0].record(stream)
        frame_resident(k)
        ev[k][1].record(stream)
        # per-stage times of this frame (synchronises after the frame's end event was recorded)
        if not args.graph:
            for mode in (N.MODE_DETECT, N.MODE_PREDICT):
                for name, ms in eng.stage_ms(mode).items():
                    key = ("detect." if mode == N.MODE_DETECT else "predict.") + name
                    stage_acc[key] = stage_acc.get(key, 0.0) + ms
        launches[0] += eng.launch_count() + (exch.launches_last if exch is not None else 0)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    lat_ms = np.array([a.elapsed_time(b) for a, b in ev], np.float64)
    t_dev = float(lat_ms.sum()) / 1e3
    counts = eng.counts()

    # ---- timed region: end to end from pinned host memory ----------------------------------------
    for k in range(min(args.warmup, 3)):
        frame_e2e(k)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    e2e_lat = []
    for k in range(args.steps):
        tf = time.perf_counter()
        pairs = frame_e2e(k)
        e2e_lat.append(time.perf_counter() - tf)
        d2h += pairs.nbytes + 96
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_serial = {"ms_per_step": t_e2e / args.steps * 1e3, "p99_ms": float(np.percentile(np.array(e2e_lat) * 1e3, 99))}
    inflight = 1
    # Two frames in flight: rcd_download_begin / _finish deliver frame k from the handle's twin pair buffer on a
    # copy stream while the kernels of frame k + 1 run.  Every frame still pays its own H2D and D2H inside the
    # timed region; a frame's latency runs from its first upload call to the arrival of its last pair.
    if args.e2e_inflight >= 2:
        pairs_pin2 = torch.empty(pairs_pin.numel(), dtype=torch.uint8).pin_memory()
        bufs = [pairs_host, pairs_pin2.numpy().view(N.PAIR_DTYPE)]

        def submit(k: int):
            p = pin[k % len(pin)]
            n = int(p["px"].shape[0])
            eng.upload_host_ptrs(n, [p[f].data_ptr() for f in FRAME_FIELDS], p["type"].data_ptr(), p["id"].data_ptr())
            eng.set_patterns_host_ptr(n, p["pattern"].data_ptr())
            if exch is not None:
                exch.exchange()
            eng.step(N.MODE_PREDICT, with_detect=True)

        def pipelined(n_frames: int):
            lat, nbytes, t_start = [], 0, {}
            t_start[0] = time.perf_counter()
            submit(0)
            eng.download_begin(bufs[0])
            for k in range(1, n_frames + 1):
                if k < n_frames:
                    t_start[k] = time.perf_counter()
                    submit(k)
                got, _c = eng.download_finish()  # frame k - 1
                lat.append(time.perf_counter() - t_start[k - 1])
                nbytes += got.nbytes + 96
                if k < n_frames:
                    eng.download_begin(bufs[k % 2])
            return lat, nbytes

        pipelined(2)  # allocate the twin buffers outside the timed region
        barrier()
        t0 = time.perf_counter()
        lat_pipe, d2h_pipe = pipelined(args.steps)
        barrier()
        t_pipe = time.perf_counter() - t0
        t_both = torch.tensor([t_pipe, t_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_both, op=dist.ReduceOp.MAX)  # every rank must take the same branch
        if float(t_both[0]) < float(t_both[1]):
            t_e2e, inflight, e2e_lat, d2h = t_pipe, 2, lat_pipe, d2h_pipe
    clocks = sampler.stop() if rank == 0 else None
    graph_replays = eng.graph_replays() if args.graph else 0
    if args.graph:  # stage breakdown on a profiled twin (outside every timed region)
        prof = FrameEngine(cap, max_pairs, device=local_rank, world_bounds=eng_bounds, profile=True)
        keep, eng = eng, prof
        ex_keep, exch = exch, (SlabExchange(prof, lo, hi, rank, world, halo,
                                            torch.cuda.ExternalStream(prof.cuda_stream(), device=torch.device("cuda", local_rank)),
                                            cap_records=int(1.1 * n_send_max) + 4096) if world > 1 else None)
        for k in range(args.steps):
            frame_resident(k)
            for mode in (N.MODE_DETECT, N.MODE_PREDICT):
                for name, ms in eng.stage_ms(mode).items():
                    key = ("detect." if mode == N.MODE_DETECT else "predict.") + name
                    stage_acc[key] = stage_acc.get(key, 0.0) + ms
        eng, exch = keep, ex_keep
        prof.close()

    # ---- reduce over ranks (max time, summed counts) -----------------------------------------------
    n_own_mean = float(np.mean([len(o["px"]) for o in own]))
    red = torch.tensor([t_dev, t_e2e, t_wall], dtype=torch.float64, device="cuda")
    sums = torch.tensor([n_own_mean, float(counts["n_pairs"]), float(counts["n_candidates"]), float(d2h) / args.steps,
                         float(counts["n_objects"] - counts["n_owned"])], dtype=torch.float64, device="cuda")
    lat_all = torch.from_numpy(lat_ms).cuda()
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(lat_all, op=dist.ReduceOp.MAX)  # a frame is done when its slowest slab is done
    t_dev, t_e2e, t_wall = (float(v) for v in red.cpu())
    objs, n_pairs, n_cand, d2h_step, n_halo = (float(v) for v in sums.cpu())
    lat = lat_all.cpu().numpy()

    if rank == 0:
        from_peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(from_peaks):
            peak, peak_src = float(json.load(open(from_peaks))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        steps = args.steps
        n_loc = n_own_mean + counts["n_objects"] - counts["n_owned"]  # rank 0: owned + halo
        stage_ms = {k: v / steps for k, v in stage_acc.items()}
        # algorithmic bytes per launch (DESIGN.md "byte model"), rank 0's objects
        npass = max(1, -(-int(np.ceil(np.log2(max(2, eng_ncells(bounds, xlo, xhi))))) // 8))
        model = {
            "keys": 102.0 * n_loc, "sort": (16.0 * npass) * n_loc, "reorder": 108.0 * n_loc,
            "pairs": 56.0 * n_loc, "narrow": 0.0, "exact": 0.0, "qorder": 0.0,
        }
        kernels = {}
        for key, ms in stage_ms.items():
            mode, name = key.split(".")
            if name in model and ms > 0:
                b = model[name]
                kernels[key] = {"ms": round(ms, 5), "alg_bytes": b, "gbs": round(b / (ms * 1e-3) / 1e9, 2) if b else None}
        dom_key = max((k for k in kernels if kernels[k]["alg_bytes"]), key=lambda k: kernels[k]["ms"])
        dom = kernels[dom_key]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom_key)
        cpu = None
        if world == 1:  # the CPU baseline is timed on rank 0 at N=1 only
            cpu_t, cpu_desc, cpu_threads = cpu_frame_seconds(frames[0], np.full(n_total, 2, np.uint8), args.cpu_budget)
            cpu = {"value": n_total / cpu_t, "unit": "object-updates/s", "cores": cpu_threads, "kind": "port",
                   "sample": cpu_desc}
        out = {
            "metric": METRIC, "value": objs * steps / t_dev, "unit": "object-updates/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": t_dev / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 pre-filter + f64 decisions", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "objects": int(objs), "objects_per_gpu": args.objects_per_gpu,
                       "frame": "ingest + index + detect-all + predict-all (performance_test.py:794-813)",
                       "partition": f"{world} x-slabs, halo {halo:.0f} m, NCCL all_to_all" if world > 1 else "single GPU",
                       "l2": "flushed with a 256 MiB write before every timed frame", "max_pairs": max_pairs,
                       "cuda_graph": bool(args.graph), "graph_replays": int(graph_replays)},
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                           "p99_reference_rule": float(np.sort(lat)[min(len(lat) - 1, int(len(lat) * 0.99))]),
                           "max": float(lat.max())},
            "frame_totals": {"pairs_emitted": n_pairs, "candidates": n_cand, "halo_objects": n_halo},
            "e2e": {"value": objs * steps / t_e2e, "unit": "object-updates/s",
                    "h2d_bytes_per_step": int(objs * H2D_BYTES_PER_OBJECT + objs * 4), "d2h_bytes_per_step": int(d2h_step),
                    "ms_per_step": t_e2e / steps * 1e3, "p99_ms": float(np.percentile(np.array(e2e_lat) * 1e3, 99)),
                    "frames_in_flight": inflight, "one_frame_in_flight": e2e_serial},
            "gpu_launches": int(launches[0]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom_key, "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "the pair kernels are fp32-ALU bound (pairs >> bytes); HBM-bound stages are in `kernels`"},
            "kernels": kernels,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
        }
        emit(json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def eng_ncells(bounds, xlo, xhi):
    cell = 100.0 * 1.002 + 0.02
    nx = int((xhi - xlo) / cell) + 1
    ny = int((bounds[1][1] - bounds[0][1]) / cell) + 1
    nz = int((bounds[1][2] - bounds[0][2]) / cell) + 1
    return nx * ny * nz


_REAL_STDOUT = None


def quiet_stdout():
    """NCCL and friends print banners to fd 1; the contract is ONE JSON line on stdout.  Everything
    but the final line goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: str):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", default="cfg4_1m_clustered3d",
                    choices=["cfg2_5k_city", "cfg3_100k_uniform2d", "cfg4_1m_clustered3d", "cfg5_10m_skew3d",
                             "cfg5_10m_skew3d_uniform_disc"])
    ap.add_argument("--objects-per-gpu", type=int, default=None)
    ap.add_argument("--max-pairs", type=int, default=32_000_000)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--graph", action="store_true",
                    help="replay rcd_step as a CUDA graph (RCD_FLAG_GRAPH): for the small, launch-bound configs")
    ap.add_argument("--e2e-inflight", type=int, default=2,
                    help="frames in flight in the end-to-end leg at N=1 (2 = result copy overlaps the next frame)")
    ap.add_argument("--host-pairs-cap", type=int, default=32_000_000,
                    help="pairs per rank the end-to-end leg copies back to (pinned) host memory per frame")
    args = ap.parse_args()
    if args.objects_per_gpu is None:
        args.objects_per_gpu = {"cfg2_5k_city": 5000, "cfg3_100k_uniform2d": 100_000, "cfg4_1m_clustered3d": PER_GPU_DEFAULT,
                                "cfg5_10m_skew3d": 1_250_000, "cfg5_10m_skew3d_uniform_disc": 1_250_000}[args.workload]
    args.warmup = max(args.warmup, 3)
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
