"""Scripted alert-lifecycle scenario shared by tests/golden/make_golden.py (reference AlertManager under the
shim), tests/test_alerts_cpu.py (oracle) and tests/test_gpu_alerts.py (device table)."""
import numpy as np


def scenario(seed: int = 11, n_vehicles: int = 60, steps: int = 14, per_step: int = 160):
    rng = np.random.default_rng(seed)
    script, now = [], 1000.0
    hot = [(int(a), int(b)) for a, b in rng.integers(0, n_vehicles, (40, 2)) if a != b]  # pairs that keep coming back
    for s in range(steps):
        now += float(rng.choice([1.5, 4.0, 9.0]))
        pairs = {}
        for _ in range(per_step):
            if hot and rng.random() < 0.5:
                i, j = hot[int(rng.integers(len(hot)))]
            else:
                i, j = (int(x) for x in rng.integers(0, n_vehicles, 2))
            if i == j:
                continue
            # fp32-representable values (the device table keeps risk / ttc as fp32), thresholds included
            risk = float(np.float32(rng.choice([0.29999, 0.3, 0.45, 0.6, 0.79, 0.8, 0.95, float(rng.random())])))
            ttc = float(np.float32(rng.choice([0.0, 2.9, 3.0, 4.9, 5.0, 7.5, float(rng.random() * 10.4)])))
            pairs[(i, j)] = (f"v{i}", f"v{j}", risk, ttc, float(np.float32(rng.random() * 10)))
        script.append(("process", now, list(pairs.values())))
        if s % 3 == 1:
            acks = [(f"v{i}", f"v{j}") for i, j in rng.integers(0, n_vehicles, (25, 2))] + \
                   [(v, o) for v, o, *_ in list(pairs.values())[:6]]
            script.append(("ack", acks))
        if s % 2 == 1:
            now += 0.25
            script.append(("cleanup", now))
    now += 31.0
    script.append(("cleanup", now))  # everything still alive is older than 30 s now
    return script
