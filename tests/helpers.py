"""Shared helpers for the parity tests (golden loading, canonical comparison)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_FIELDS = ("px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az", "size", "heading", "type")

# north_star tolerance: pair sets bit-exact; TTC / distances within 1e-4 relative (fp32 vs f64)
RTOL = 1e-4


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    frame = {k: z[f"frame_{k}"] for k in FRAME_FIELDS if f"frame_{k}" in z.files}
    return z, frame


def f64_frame(frame):
    out = {k: np.asarray(v, np.float64) for k, v in frame.items() if k != "type"}
    out["type"] = np.asarray(frame["type"], np.int32)
    return out


def oracle_risk_table(risks):
    """structured oracle risks -> (n, 9) float table in golden column order."""
    cols = ("i", "j", "ttc", "distance", "rel_speed", "risk", "cx", "cy", "cz")
    return np.stack([risks[c].astype(np.float64) for c in cols], 1) if len(risks) else np.zeros((0, 9))


def assert_pairs_equal(got_ij, want_ij, what="pairs"):
    got = np.asarray(got_ij, np.int64).reshape(-1, 2)
    want = np.asarray(want_ij, np.int64).reshape(-1, 2)
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    want = want[np.lexsort((want[:, 1], want[:, 0]))]
    if got.shape != want.shape or not np.array_equal(got, want):
        gs = set(map(tuple, got.tolist()))
        ws = set(map(tuple, want.tolist()))
        raise AssertionError(f"{what}: {len(gs)} got vs {len(ws)} wanted; missing {sorted(ws - gs)[:8]} "
                             f"extra {sorted(gs - ws)[:8]}")


def assert_close(got, want, what, rtol=RTOL, atol=0.0):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + atol
    if np.any(err > tol):
        k = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: max violation at {k}: got {got.flat[k]!r} want {want.flat[k]!r} "
                             f"(rtol {rtol}, atol {atol})")
