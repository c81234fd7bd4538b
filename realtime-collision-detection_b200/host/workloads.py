"""Synthetic frames for the BASELINE.json configs.

A *frame* is a dict of 1-D numpy arrays, one entry per object, in the SoA layout the C-ABI
takes (include/rcd.h, ``rcd_upload``): ``px py pz vx vy vz ax ay az size heading`` float32 and
``type`` uint8.  All values are fp32, i.e. the frame *is* its fp32 representation: the oracle is
fed exactly these values widened to float64, so both sides see identical inputs.

Distributions follow the reference's own generators:
  * ``reference_city_frame``  -- src/test/performance_test.py:52-145 (5 "cities", 80 % inside a
    disc with r = U*radius, 20 % uniform; speed U(5,20); heading U(0, 2*3.14159); size by type)
    and the per-step motion update of :147-195 (``advance``).
  * drones -- src/test/load_generator.py:473-503 (z U(0,100), vz U(-5,5), az U(-1,1), size U(1,5)).
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Tuple

import numpy as np

Frame = Dict[str, np.ndarray]
FRAME_FIELDS = ("px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az", "size", "heading")

TWO_PI_REF = 2 * 3.14159  # the reference uses this literal, not math.pi
TYPE_SIZES = np.array([2.0, 4.0, 5.0, 1.0], np.float32)  # car, truck, bus, motorcycle
TYPE_DRONE = 4


def empty_frame(n: int = 0) -> Frame:
    f = {k: np.zeros(n, np.float32) for k in FRAME_FIELDS}
    f["type"] = np.zeros(n, np.uint8)
    return f


def _finish(f: Dict[str, np.ndarray]) -> Frame:
    out = {k: np.ascontiguousarray(f[k], dtype=np.float32) for k in FRAME_FIELDS}
    out["type"] = np.ascontiguousarray(f["type"], dtype=np.uint8)
    return out


def frame_to_f64(frame: Frame) -> Dict[str, np.ndarray]:
    """The same values widened to float64 (what the oracle consumes)."""
    out = {k: frame[k].astype(np.float64) for k in FRAME_FIELDS}
    out["type"] = frame["type"].astype(np.int32)
    return out


def take(frame: Frame, idx) -> Frame:
    return {k: np.ascontiguousarray(v[idx]) for k, v in frame.items()}


def concat(a: Frame, b: Frame) -> Frame:
    return {k: np.concatenate([a[k], b[k]]) for k in a}


# ----------------------------------------------------------------------------------------
# configs[0] / configs[1]: the reference perf-test generator, statement by statement
# ----------------------------------------------------------------------------------------
def reference_city_frame(n: int, seed: int, map_size: Tuple[float, float] = (10000.0, 10000.0)) -> Frame:
    """performance_test.py:82-145 with Python's ``random`` seeded (the reference is unseeded)."""
    rng = random.Random(seed)
    w, h = map_size
    cities = [(w * 0.25, h * 0.25, 1000.0), (w * 0.75, h * 0.25, 1000.0), (w * 0.25, h * 0.75, 1000.0),
              (w * 0.75, h * 0.75, 1000.0), (w * 0.5, h * 0.5, 2000.0)]
    f = {k: np.zeros(n, np.float64) for k in FRAME_FIELDS}
    typ = np.zeros(n, np.uint8)
    for i in range(n):
        t = rng.randrange(4)
        typ[i] = t
        f["size"][i] = float(TYPE_SIZES[t])
        if rng.random() < 0.8:
            cx, cy, cr = cities[rng.randrange(5)]
            r = rng.random() * cr
            th = rng.random() * TWO_PI_REF
            f["px"][i] = cx + r * np.cos(th)
            f["py"][i] = cy + r * np.sin(th)
        else:
            f["px"][i] = rng.uniform(0, w)
            f["py"][i] = rng.uniform(0, h)
        speed = rng.uniform(5, 20)
        hd = rng.uniform(0, TWO_PI_REF)
        f["vx"][i] = speed * np.cos(hd)
        f["vy"][i] = speed * np.sin(hd)
        f["heading"][i] = hd
    f["type"] = typ
    return _finish(f)


def advance(frame: Frame, dt: float, rng: np.random.Generator,
            map_size: Tuple[float, float] = (10000.0, 10000.0), max_speed: float = 30.0) -> Frame:
    """One motion step of the reference generator (performance_test.py:147-195), vectorised:
    integrate, reflect at the borders, 10 % chance of a new U(-1,1) xy acceleration, speed cap,
    heading = atan2(vy, vx) when moving.  Drones (type 4) also integrate z and stay in [0, 100]."""
    f = {k: frame[k].astype(np.float64) for k in FRAME_FIELDS}
    n = f["px"].shape[0]
    for p, v, lim in (("px", "vx", map_size[0]), ("py", "vy", map_size[1])):
        f[p] += f[v] * dt
        lo = f[p] < 0
        hi = f[p] > lim
        f[p][lo] = 0.0
        f[p][hi] = lim
        f[v][lo | hi] *= -1.0
    f["pz"] += f["vz"] * dt
    lo = f["pz"] < 0
    hi = f["pz"] > 100.0
    f["pz"][lo] = 0.0
    f["pz"][hi] = 100.0
    f["vz"][lo | hi] *= -1.0
    change = rng.random(n) < 0.1
    k = int(change.sum())
    f["ax"][change] = rng.uniform(-1, 1, k)
    f["ay"][change] = rng.uniform(-1, 1, k)
    f["vx"] += f["ax"] * dt
    f["vy"] += f["ay"] * dt
    speed = np.sqrt(f["vx"] ** 2 + f["vy"] ** 2)
    fast = speed > max_speed
    f["vx"][fast] = f["vx"][fast] / speed[fast] * max_speed
    f["vy"][fast] = f["vy"][fast] / speed[fast] * max_speed
    moving = speed > 0.1
    f["heading"][moving] = np.arctan2(f["vy"][moving], f["vx"][moving])
    f["type"] = frame["type"]
    return _finish(f)


# ----------------------------------------------------------------------------------------
# vectorised generators for the large configs
# ----------------------------------------------------------------------------------------
def _ground_kinematics(n: int, rng: np.random.Generator, f: Dict[str, np.ndarray], accel: bool) -> None:
    typ = rng.integers(0, 4, n).astype(np.uint8)
    f["type"] = typ
    f["size"] = TYPE_SIZES[typ].astype(np.float64)
    speed = rng.uniform(5, 20, n)
    hd = rng.uniform(0, TWO_PI_REF, n)
    f["vx"] = speed * np.cos(hd)
    f["vy"] = speed * np.sin(hd)
    f["vz"] = np.zeros(n)
    f["heading"] = hd
    if accel:  # steady state of the reference motion model: U(-1,1) xy accelerations
        f["ax"] = rng.uniform(-1, 1, n)
        f["ay"] = rng.uniform(-1, 1, n)
    else:
        f["ax"] = np.zeros(n)
        f["ay"] = np.zeros(n)
    f["az"] = np.zeros(n)
    f["pz"] = np.zeros(n)


def _make_drones(mask: np.ndarray, rng: np.random.Generator, f: Dict[str, np.ndarray]) -> None:
    k = int(mask.sum())
    f["type"][mask] = TYPE_DRONE
    f["pz"][mask] = rng.uniform(0, 100, k)
    f["vz"][mask] = rng.uniform(-5, 5, k)
    f["az"][mask] = rng.uniform(-1, 1, k)
    f["size"][mask] = rng.uniform(1, 5, k)


def uniform_frame(n: int, seed: int, map_size: float = 10000.0, accel: bool = True,
                  drone_fraction: float = 0.0) -> Frame:
    """configs[2]: n objects uniform on a map_size x map_size map, z = 0 (2-D)."""
    rng = np.random.default_rng(seed)
    f: Dict[str, np.ndarray] = {}
    _ground_kinematics(n, rng, f, accel)
    f["px"] = rng.uniform(0, map_size, n)
    f["py"] = rng.uniform(0, map_size, n)
    if drone_fraction > 0:
        _make_drones(rng.random(n) < drone_fraction, rng, f)
    return _finish(f)


def hotspot_frame(n: int, seed: int, map_size: float, n_hotspots: int, zipf_s: float = 0.0,
                  hotspot_fraction: float = 0.8, radius_range: Tuple[float, float] = (1000.0, 2000.0),
                  drone_fraction: float = 0.3, radial_law: str = "reference", accel: bool = True) -> Frame:
    """configs[3] (1 M, 50 equal hotspots) and configs[4] (10 M, 200 hotspots, Zipf(s=1) weights).

    radial_law 'reference': r = U * radius  (density ~ 1/r, performance_test.py:92-103);
               'uniform'  : r = sqrt(U) * radius (uniform in the disc; the feasibility variant of
                            SURVEY.md 8d config 5).
    """
    rng = np.random.default_rng(seed)
    f: Dict[str, np.ndarray] = {}
    _ground_kinematics(n, rng, f, accel)
    centres = rng.uniform(0, map_size, (n_hotspots, 2))
    radii = rng.uniform(radius_range[0], radius_range[1], n_hotspots)
    w = 1.0 / np.arange(1, n_hotspots + 1) ** zipf_s if zipf_s > 0 else np.ones(n_hotspots)
    w = w / w.sum()
    in_hot = rng.random(n) < hotspot_fraction
    which = rng.choice(n_hotspots, size=n, p=w)
    u = rng.random(n)
    r = (u if radial_law == "reference" else np.sqrt(u)) * radii[which]
    th = rng.random(n) * TWO_PI_REF
    hx = centres[which, 0] + r * np.cos(th)
    hy = centres[which, 1] + r * np.sin(th)
    f["px"] = np.where(in_hot, hx, rng.uniform(0, map_size, n))
    f["py"] = np.where(in_hot, hy, rng.uniform(0, map_size, n))
    if drone_fraction > 0:
        _make_drones(rng.random(n) < drone_fraction, rng, f)
    return _finish(f)


def random_patterns(n: int, seed: int, p=(0.1, 0.5, 0.3, 0.1)) -> np.ndarray:
    """Per-object trajectory-pattern codes (0 stationary, 1 constant_velocity, 2 accelerating,
    3 no history -> detect path), used by predict-mode tests."""
    return np.random.default_rng(seed).choice(4, size=n, p=p).astype(np.uint8)


WORKLOADS = {
    # name: (description, factory(n_override) -> Frame)
    "cfg1_1k_city": ("reference perf test: 1000 vehicles (configs[0])",
                     lambda n=None: reference_city_frame(n or 1000, 1234)),
    "cfg2_5k_city": ("reference perf test: 5000 vehicles (configs[1])",
                     lambda n=None: reference_city_frame(n or 5000, 1235)),
    "cfg3_100k_uniform2d": ("100k uniform vehicles, 2-D, 10 km map (configs[2])",
                            lambda n=None: uniform_frame(n or 100_000, 2001)),
    "cfg4_1m_clustered3d": ("1M vehicles+drones, 3-D, 50 hotspots, 31.6 km map (configs[3])",
                            lambda n=None: hotspot_frame(n or 1_000_000, 2002, 31623.0, 50)),
    "cfg5_10m_skew3d": ("10M objects, 3-D, 200 Zipf hotspots, 100 km map (configs[4])",
                        lambda n=None: hotspot_frame(n or 10_000_000, 2003, 100000.0, 200, zipf_s=1.0)),
}


def make_workload(name: str, n: Optional[int] = None) -> Frame:
    return WORKLOADS[name][1](n)
