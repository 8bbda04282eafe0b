"""The C restatement (oracle/dlrm_oracle.c) against the golden-pinned numpy oracle."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import oracle as O
from tests.helpers import golden_model, load_golden


@pytest.mark.parametrize("B,F,d,pad", [(128, 8, 16, 1), (50, 27, 64, 1), (9, 11, 128, 8), (3, 6, 10, 1), (4, 1, 8, 1)])
def test_c_interaction_matches_numpy(B, F, d, pad):
    rng = np.random.default_rng(B + F + d)
    T = rng.standard_normal((B, F, d)).astype(np.float32)
    out = CO.interaction_fwd(T, pad)
    assert O.rel_err(out, O.interaction_fwd(T, pad)) < 1e-6
    g = rng.standard_normal(out.shape).astype(np.float32)
    dx, dT = CO.interaction_bwd(g, T, pad)
    dx_ref, dT_ref = O.interaction_bwd(g, T, out.shape[1] - d - F * (F - 1) // 2)
    assert O.rel_err(dx, dx_ref) < 1e-6
    assert O.rel_err(dT, dT_ref) < 1e-6 or F == 1


@pytest.mark.parametrize("name", ["single", "multi"])
def test_c_lookup_and_update_on_goldens(name):
    g = load_golden(name)
    _bot, _top, tables, _dense, idx, _labels = golden_model(g)
    T = CO.lookup(tables, idx, slot0=1)
    assert np.array_equal(T, O.lookup(tables, idx, slot0=1))
    assert O.rel_err(T[:, 1:], g["concatenated_result"][:, 1:]) < 1e-6
    out = CO.interaction_fwd(g["concatenated_result"])
    assert O.rel_err(out, g["output_interaction"]) < 1e-6
    rng = np.random.default_rng(0)
    dT = rng.standard_normal(T.shape).astype(np.float32)
    a = [t.copy() for t in tables]
    b = [t.copy() for t in tables]
    CO.sparse_sgd(a, idx, dT, 1, 10.0)
    for k in range(7):
        O.sparse_sgd_update_fast(b[k], idx[k], np.ascontiguousarray(dT[:, 1 + k]), 10.0)
        assert O.rel_err(a[k], b[k]) < 1e-6


def test_c_oracle_thread_count_invariance():
    rng = np.random.default_rng(1)
    T = rng.standard_normal((64, 27, 64)).astype(np.float32)
    assert np.array_equal(CO.interaction_fwd(T, nthreads=1), CO.interaction_fwd(T, nthreads=4))
