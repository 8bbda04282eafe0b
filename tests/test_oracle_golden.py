"""Pin the CPU oracle against every known-answer vector the reference holds for the path.

Mirrors the reference's own ladder: 3x3 integer triangle
(test/model/interact.jl:9-32) -> random small vs naive reference (:36-58) -> hand-typed
4-sample PyTorch case (test/model/model.jl:80-283) -> both ref/*.hdf5 goldens
(test/integration.jl, src/validation.jl).
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import golden_model, golden_updates, load_golden, load_known_answer


def test_triangle_3x3_integer_examples():
    x = np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]])
    y = O.triangular_slice_kernel(x)
    assert y.tolist() == [4, 7, 8]
    assert O.triangular_slice_back_kernel(y, 3).tolist() == [[0, 4, 7], [0, 0, 8], [0, 0, 0]]
    assert O.triangular_slice_back_fuse_add_transpose_kernel(y, 3).tolist() == [[0, 4, 7], [4, 0, 8], [7, 8, 0]]


@pytest.mark.parametrize("n", list(range(5, 25)))
def test_triangle_random_matches_reference_walk(n):
    rng = np.random.default_rng(n)
    x = rng.random((n, n), dtype=np.float32)
    y = O.triangular_slice_kernel(x)
    # triangular_slice_reference (interact.jl:26-31): vcat of x[1:i-1, i] for i = 2..n
    ref = np.concatenate([x[:i, i] for i in range(1, n)])
    assert np.array_equal(y, ref)
    up = O.triangular_slice_back_kernel(y, n)
    assert np.array_equal(up, np.triu(x, 1))
    assert np.array_equal(O.triangular_slice_back_fuse_add_transpose_kernel(y, n), up + up.T)


def test_flat_pair_position_formula():
    # pair i<j (0-based) sits at j(j-1)/2 + i; G symmetric so walking the
    # strict upper triangle column-major == strict lower triangle row-major.
    F, d = 9, 4
    rng = np.random.default_rng(0)
    T = rng.standard_normal((3, F, d)).astype(np.float32)
    out = O.interaction_fwd(T)
    for b in range(3):
        G = T[b] @ T[b].T
        assert np.allclose(out[b, d:], O.triangular_slice_kernel(G), rtol=1e-6, atol=1e-6)
    assert np.allclose(out, O.interaction_fwd_loops(T), rtol=1e-6, atol=1e-6)


def test_interaction_backward_is_the_gradient():
    # finite-difference check in float64 of the restated pullback (interact.jl:424-436)
    rng = np.random.default_rng(1)
    B, F, d = 2, 5, 4
    T = rng.standard_normal((B, F, d))
    g = rng.standard_normal((B, d + O.num_pairs(F)))
    dx, dT = O.interaction_bwd(g.astype(np.float32), T.astype(np.float32))

    def f(Tm):
        jj, ii = np.tril_indices(F, -1)
        G = np.einsum("bik,bjk->bij", Tm, Tm)
        return (g[:, :d] * Tm[:, 0]).sum() + (g[:, d:] * G[:, jj, ii]).sum()

    num = np.zeros_like(T)
    h = 1e-6
    for i in np.ndindex(*T.shape):
        Tp = T.copy(); Tp[i] += h
        Tm = T.copy(); Tm[i] -= h
        num[i] = (f(Tp) - f(Tm)) / (2 * h)
    # total gradient wrt slot 0 is dx; wrt other slots dT
    assert np.allclose(dx, num[:, 0], rtol=1e-4, atol=1e-4)
    assert np.allclose(dT[:, 1:], num[:, 1:], rtol=1e-4, atol=1e-4)


def test_known_answer_small_pytorch_case():
    ka = load_known_answer()
    tables = [ka[f"py_embedding{i}_weights"] for i in (1, 2, 3)]
    idx = [np.asarray(v) for v in ka["py_sparse_input"]]
    bot = [(ka["py_dense1_weights"], ka["py_dense1_bias"])]
    top = [(ka["py_dense2_weights"], ka["py_dense2_bias"]),
           (ka["py_dense3_weights"].reshape(1, -1), ka["py_dense3_bias"])]
    fwd = O.dlrm_forward(bot, top, tables, ka["py_dense_input"], idx)
    # constants are typed to 5 decimals -> absolute gate 1e-5 (+ half-ulp of the rounding)
    assert np.allclose(fwd["x"], ka["py_bottom_mlp_output"], atol=1.5e-5)
    for k in range(3):
        assert np.allclose(fwd["T"][:, 1 + k], np.asarray(ka["py_embedding_outputs"][k], dtype=np.float32), atol=1e-6)
    assert np.allclose(fwd["z"], ka["py_interaction_output"], atol=1.5e-5)
    assert np.allclose(fwd["out"], ka["py_top_mlp_output"], atol=1.5e-5)


@pytest.mark.parametrize("name", ["single", "multi"])
def test_hdf5_golden_forward(name):
    g = load_golden(name)
    bot, top, tables, dense, idx, labels = golden_model(g)
    fwd = O.dlrm_forward(bot, top, tables, dense, idx)
    assert O.rel_err(fwd["x"], g["mlp_bottom"]) < 1e-6
    assert O.rel_err(fwd["T"], g["concatenated_result"]) < 1e-6      # lookup + sum pool
    assert np.array_equal(fwd["T"][:, 1:], g["concatenated_result"][:, 1:]) or name == "multi"
    jj, ii = np.tril_indices(8, -1)
    G = np.einsum("bik,bjk->bij", fwd["T"], fwd["T"])
    assert O.rel_err(G, g["zpre"]) < 1e-6
    assert O.rel_err(fwd["z"][:, 16:], g["zflat"]) < 1e-6
    assert O.rel_err(fwd["z"], g["output_interaction"]) < 1e-6
    assert O.rel_err(fwd["out"], g["mlp_top"].reshape(-1)) < 1e-6
    assert abs(float(O.bce_loss(fwd["out"], labels)) - float(g["loss"])) < 1e-6


@pytest.mark.parametrize("name", ["single", "multi"])
def test_hdf5_golden_one_sgd_step(name):
    """src/validation.jl:1-44: lr = 10.0, one step; every table and MLP must match."""
    g = load_golden(name)
    bot, top, tables, dense, idx, labels = golden_model(g)
    orig = [t.copy() for t in tables]
    loss, nbot, ntop, _fwd, _gr = O.dlrm_train_step(bot, top, tables, dense, idx, labels, lr=10.0)
    ubot, utop, uemb = golden_updates(g)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for k in range(7):
        assert O.rel_err(tables[k], uemb[k]) < 1e-6, k
        assert not O.isapprox(orig[k], uemb[k]), "golden update must differ from original"
    for (W, b), (uW, ub) in zip(nbot + ntop, ubot + utop):
        assert O.rel_err(W, uW) < 1e-5
        assert O.rel_err(b, ub) < 1e-5


def test_sparse_update_loop_and_fast_agree_bitwise():
    rng = np.random.default_rng(3)
    for P in (1, 3):
        table = rng.standard_normal((50, 8)).astype(np.float32)
        idx = rng.integers(0, 50, size=(64, P))
        delta = rng.standard_normal((64, 8)).astype(np.float32)
        a, b = table.copy(), table.copy()
        O.sparse_sgd_update(a, idx, delta, 0.1)
        O.sparse_sgd_update_fast(b, idx, delta, 0.1)
        assert np.array_equal(a, b)
        dense = O.uncompress(delta, idx, 50)
        assert O.rel_err(a, table - np.float32(0.1) * dense) < 1e-6


def test_sort_dedup_definition():
    idx = np.array([5, 1, 5, 0, 1, 5])
    uniq, seg, perm = O.sort_dedup(idx)
    assert uniq.tolist() == [0, 1, 5]
    assert seg.tolist() == [0, 1, 3, 6]
    assert perm.tolist() == [3, 1, 4, 0, 2, 5]
