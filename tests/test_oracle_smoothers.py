"""CPU: the MULTICOLOR_GS restatement of the oracle against a textbook numpy multicolour Gauss-Seidel (parity unpinned: no
reference golden yet; tests/golden/make_golden.py lists the case to generate)."""
import numpy as np
import pytest

from amgx_b200 import gallery


def sym_banded(n, sigma):
    """structurally symmetric (a proper colouring needs i ~ j <=> j ~ i), diagonally dominant"""
    rp, ci, va = gallery.random_banded(n, sigma=sigma)
    A = gallery.to_scipy(rp, ci, va)
    A = (A + A.T).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def numpy_gs(A, colors, nc, b, x, w, symmetric):
    A = A.tocsr()
    d = A.diagonal()
    x = x.copy()
    order = list(range(nc)) + (list(range(nc - 1, -1, -1)) if symmetric else [])
    for c in order:
        rows = np.nonzero(colors == c)[0]
        r = b[rows] - (A[rows] @ x)
        x[rows] = x[rows] + w * r / d[rows]
    return x


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("mat", ["poisson", "banded"])
def test_gs_sweep_matches_textbook(oracle, sym, mat):
    rp, ci, va = gallery.poisson7pt(9, 7, 5) if mat == "poisson" else sym_banded(800, 12.0)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    rng = np.random.default_rng(3)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    nc, colors, srows, offs = oracle.color_min_max(rp, ci, 0.0)
    # a proper colouring: no two coupled rows share a colour
    coo = A.tocoo()
    off = coo.row != coo.col
    assert not np.any(colors[coo.row[off]] == colors[coo.col[off]])
    x = oracle.gs_sweep(rp, ci, va, b, x0, 0.9, symmetric=sym)
    ref = numpy_gs(A, colors, nc, b, x0, 0.9, sym)
    assert np.max(np.abs(x - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_gs_is_exact_for_a_diagonal_matrix_and_reduces_the_error(oracle):
    n = 50
    rp = np.arange(n + 1, dtype=np.int32)
    ci = np.arange(n, dtype=np.int32)
    va = np.linspace(1.0, 3.0, n)
    b = np.ones(n)
    x = oracle.gs_sweep(rp, ci, va, b, np.zeros(n), 1.0)
    assert np.allclose(x, b / va, rtol=1e-15)
    rp, ci, va = gallery.poisson7pt(8)
    A = gallery.to_scipy(rp, ci, va)
    N = rp.shape[0] - 1
    xt = np.random.default_rng(0).standard_normal(N)
    b = A @ xt
    x = np.zeros(N)
    e0 = np.sqrt((xt - x) @ (A @ (xt - x)))
    for _ in range(5):
        x = oracle.gs_sweep(rp, ci, va, b, x, 1.0, symmetric=True)
    e1 = np.sqrt((xt - x) @ (A @ (xt - x)))
    assert e1 < 0.5 * e0


@pytest.mark.parametrize("sym", [False, True])
def test_amg_with_gs_smoother_converges(oracle, sym):
    oracle.set_uncolored_fraction(0.0)
    try:
        rp, ci, va = gallery.poisson7pt(14)
        n = rp.shape[0] - 1
        amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.9, smoother="MULTICOLOR_GS").set_symmetric_gs(sym)
        x, it, hist, conv = oracle.amg_solve(amg, np.ones(n), tol=1e-8, max_iters=80)
        jac = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.9)
        xj, itj, histj, convj = oracle.amg_solve(jac, np.ones(n), tol=1e-8, max_iters=80)
        assert conv and convj and it < itj
        A = gallery.to_scipy(rp, ci, va)
        assert np.linalg.norm(np.ones(n) - A @ x) <= 1.01e-8 * np.sqrt(n)
    finally:
        oracle.set_uncolored_fraction(0.15)
