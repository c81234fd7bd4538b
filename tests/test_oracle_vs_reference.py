"""Build-container-only: the oracle against the reference's own bytecode (shim) on FRESH seeds,
so the pin is not limited to the committed fixtures.  Skipped where /root/reference is absent."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import oracle_risk_table

pytestmark = pytest.mark.needs_reference


def _frame(seed, n=300, box=350.0):
    from rcd_b200.host import workloads as W
    return W.frame_to_f64(W.uniform_frame(n, seed, map_size=box, drone_fraction=0.3))


@pytest.mark.parametrize("seed", [1, 2])
def test_detect_fresh_seed(seed):
    from oracle import ref_shim as S
    f = _frame(seed)
    ref = S.run_detect_A(f)
    o = O.frame_A(f, "detect", want_candidates=True)
    assert np.array_equal(o["candidates"], np.array(ref["candidates"], np.int32).reshape(-1, 2))
    want = np.array([[float(x) for x in r] for r in ref["risks"]]).reshape(-1, 9)
    assert np.array_equal(oracle_risk_table(o["risks"]), want)


def test_predict_fresh_seed():
    from oracle import ref_shim as S
    from rcd_b200.host import workloads as W
    f = _frame(3, n=200, box=300.0)
    pat = W.random_patterns(200, 4)
    ref = S.run_predict_A(f, pat)
    o = O.frame_A(f, "predict", pattern_codes=pat)
    want = np.array([[float(x) for x in r[:9]] for r in ref["risks"]]).reshape(-1, 9)
    assert len(want) > 20
    assert np.array_equal(oracle_risk_table(o["risks"]), want)
