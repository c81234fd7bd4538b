"""The oracle's alert table (oracle/oracle.py AlertTable) against the reference's AlertManager on the
scripted scenario (tests/golden/alert_scenario.json.gz: reference bytecode under the shim, controlled
clock).  No GPU needed; the device table is compared with the oracle in tests/test_gpu_alerts.py."""
import gzip
import json
import os

import pytest

from tests import alert_cases as A

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "alert_scenario.json.gz")


def load_golden_alerts():
    with gzip.open(GOLDEN) as f:
        return json.loads(f.read().decode())


def run_oracle(script):
    from oracle import oracle as O
    t = O.AlertTable()
    out = []
    for op in script:
        events = []
        if op[0] == "process":
            events = t.process([(v, o, r, ttc) for v, o, r, ttc, _d in op[2]], op[1])
        elif op[0] == "ack":
            t.acknowledge(op[1])
        else:
            events = [("expired", v, o, -1, -1) for v, o in t.cleanup(op[1])]
        table = sorted((k[0], k[1], a["risk"], a["ttc"], a["priority"], a["timestamp"], a["acknowledged"])
                       for k, a in t.alerts.items())
        out.append((sorted(events), table))
    return out


def test_oracle_alert_table_matches_reference_golden():
    rows = load_golden_alerts()
    got = run_oracle(A.scenario())
    assert len(rows) == len(got)
    kinds = set()
    for row, (events, table) in zip(rows, got):
        assert [list(e) for e in events] == row["events"]
        want = [(t[0], t[1], float(t[2]), float(t[3]), t[4], float(t[5]), t[6]) for t in row["table"]]
        assert table == want
        kinds |= {e[0] for e in events}
    assert kinds == {"created", "changed", "refreshed", "expired"}


@pytest.mark.needs_reference
def test_oracle_alert_table_matches_live_reference_on_a_fresh_seed():
    from oracle import ref_shim as S
    script = A.scenario(seed=99, n_vehicles=30, steps=8, per_step=90)
    ref = S.run_alert_scenario_A(script)
    assert ref == run_oracle(script)
