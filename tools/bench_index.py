"""Index-build microbenchmark (keys -> onesweep sort -> reorder) at sizes where the kernels are
HBM-bound rather than launch-bound.  Prints one JSON line per size with achieved GB/s against the
algorithmic byte model of DESIGN.md section 5."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from rcd_b200.host import _native as N
from rcd_b200.host.engine import FrameEngine, FRAME_FIELDS

peak = 6546.2
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
sizes = [int(s) for s in sys.argv[1:]] or [1_000_000, 8_000_000, 32_000_000]
for n in sizes:
    side = float(np.sqrt(n / 1e-3))  # 1e-3 objects / m^2, about 10 per 100 m cell
    g = torch.Generator(device="cuda").manual_seed(1)
    d = {k: torch.zeros(n, dtype=torch.float32, device="cuda") for k in FRAME_FIELDS}
    d["px"] = torch.rand(n, generator=g, device="cuda") * side
    d["py"] = torch.rand(n, generator=g, device="cuda") * side
    d["vx"] = torch.randn(n, generator=g, device="cuda") * 10
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with FrameEngine(n, 1024, world_bounds=((0, 0, 0), (side, side, 0)), profile=True) as e:
        e.upload_device(n, [d[k].data_ptr() for k in FRAME_FIELDS])
        best = None
        for r in range(8):
            flush.zero_(); torch.cuda.synchronize()
            e.invalidate()
            e.build_index(100.0)
            ms = e.stage_ms(N.MODE_DETECT)
            if r >= 2 and (best is None or ms["total"] < best["total"]):
                best = ms
        ncells = (int(side / 100.22) + 1) ** 2
        passes = max(1, -(-int(np.ceil(np.log2(ncells))) // 8))
        # K1 50N + 4N + 52N; K2 16N per pass, 12N for the first (no permutation read); K3 4N + 52N + 56N
        model = {"keys": 106.0 * n, "sort": (16.0 * passes - 4.0) * n, "reorder": 112.0 * n}
        out = {"n": n, "passes": passes, "peak_gbs": peak}
        for k, b in model.items():
            out[k] = {"ms": round(best[k], 4), "alg_bytes": b, "gbs": round(b / best[k] / 1e6, 1), "frac": round(b / best[k] / 1e6 / peak, 3)}
        out["total_ms"] = round(best["total"], 4)
        print(json.dumps(out), flush=True)
    del d
    torch.cuda.empty_cache()
