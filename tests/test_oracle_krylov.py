"""CPU: the CG / PCGF / PBICGSTAB / GMRES restatements of the oracle (oracle/krylov_oracle.inc.c) against independent numpy
textbook formulations and against each other.  No reference golden exists for these four yet (parity unpinned; the cases
tests/golden/make_golden.py would generate are listed there) -- these tests pin the algebra, not the reference's rounding."""
import numpy as np
import pytest

from amgx_b200 import gallery


def _sys(n=9, seed=0):
    rp, ci, va = gallery.poisson7pt(n)
    N = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    b = np.random.default_rng(seed).standard_normal(N)
    return rp, ci, va, A, b


def test_cg_matches_textbook(oracle):
    rp, ci, va, A, b = _sys()
    x, it, hist, conv = oracle.krylov("CG", rp, ci, va, b, tol=1e-10, max_iters=200)
    assert conv
    # textbook CG
    xs = np.zeros_like(b); r = b.copy(); p = r.copy(); rr = r @ r; h = [np.linalg.norm(r)]
    for _ in range(it):
        Ap = A @ p; al = rr / (Ap @ p); xs += al * p; r -= al * Ap; h.append(np.linalg.norm(r))
        rn = r @ r; p = r + (rn / rr) * p; rr = rn
    assert np.allclose(hist, h, rtol=1e-9)
    assert np.allclose(x, xs, rtol=1e-9, atol=1e-12)
    assert hist[-1] <= 1e-10 * hist[0] < hist[-2]


def test_cg_equals_unpreconditioned_pcg(oracle):
    """CG and PCG without a preconditioner are the same recurrence (z = r): identical histories bit for bit"""
    rp, ci, va, A, b = _sys(8, 3)
    x1, it1, h1, c1 = oracle.krylov("CG", rp, ci, va, b, tol=1e-9, max_iters=150)
    x2, it2, h2, c2 = oracle.pcg(rp, ci, va, b, tol=1e-9, max_iters=150)
    assert it1 == it2 and c1 and c2
    assert np.allclose(h1, h2, rtol=1e-12)


@pytest.mark.parametrize("precond", ["none", "jacobi", "amg"])
def test_pcgf_equals_pcg_for_a_fixed_preconditioner(oracle, precond):
    """for a fixed linear SPD preconditioner <z_new, r_new - r_old> = <z_new, r_new> in exact arithmetic: same iterates as PCG"""
    rp, ci, va, A, b = _sys(10, 1)
    kw = {}
    if precond == "jacobi":
        kw["jacobi_omega"] = 0.8
    if precond == "amg":
        kw["amg"] = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8)
    x1, it1, h1, c1 = oracle.krylov("PCGF", rp, ci, va, b, tol=1e-9, max_iters=150, **kw)
    x2, it2, h2, c2 = oracle.pcg(rp, ci, va, b, tol=1e-9, max_iters=150, **kw)
    assert c1 and c2 and abs(it1 - it2) <= 1
    m = min(len(h1), len(h2))
    assert np.allclose(h1[:m], h2[:m], rtol=1e-6)
    assert np.linalg.norm(b - A @ x1) <= 1.01e-9 * np.linalg.norm(b)


@pytest.mark.parametrize("precond", ["none", "jacobi", "amg"])
def test_pbicgstab_matches_textbook(oracle, precond):
    rp, ci, va, A, b = _sys(9, 2)
    kw = {}
    M = lambda v: v.copy()
    if precond == "jacobi":
        kw["jacobi_omega"] = 0.8
        d = A.diagonal()
        M = lambda v: 0.8 * v / d
    if precond == "amg":
        amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8)
        kw["amg"] = amg
        M = lambda v: amg.vcycle(v)
    x, it, hist, conv = oracle.krylov("PBICGSTAB", rp, ci, va, b, tol=1e-9, max_iters=100, **kw)
    assert conv and np.linalg.norm(b - A @ x) <= 1.5e-9 * np.linalg.norm(b)
    # van der Vorst's preconditioned BiCGStab
    xs = np.zeros_like(b); r = b.copy(); rt = r.copy(); rho = rt @ r; p = r.copy(); h = [np.linalg.norm(r)]
    for k in range(it):
        Mp = M(p); v = A @ Mp; al = rho / (rt @ v); s = r - al * v
        if np.linalg.norm(s) <= 1e-9 * h[0]:
            xs += al * Mp; h.append(np.linalg.norm(b - A @ xs)); break
        Ms = M(s); t = A @ Ms; om = (t @ s) / (t @ t)
        xs += al * Mp + om * Ms; r = s - om * t; h.append(np.linalg.norm(r))
        rn = rt @ r; be = (rn / rho) * (al / om); rho = rn; p = r + be * p - be * om * v
    assert len(h) == len(hist)
    assert np.allclose(hist, h, rtol=1e-6)


@pytest.mark.parametrize("restart", [4, 30])
def test_gmres_equals_fgmres_without_preconditioner(oracle, restart):
    """GMRES and FGMRES build the same Krylov space when M = I: same residual estimates, same x"""
    rp, ci, va, A, b = _sys(8, 5)
    x1, it1, h1, c1 = oracle.krylov("GMRES", rp, ci, va, b, tol=1e-9, max_iters=120, restart=restart)
    x2, it2, h2, c2 = oracle.fgmres(rp, ci, va, b, tol=1e-9, max_iters=120, restart=restart)
    assert c1 and c2 and it1 == it2
    assert np.allclose(h1, h2, rtol=1e-8)
    assert np.allclose(x1, x2, rtol=1e-8, atol=1e-12)


def test_gmres_right_preconditioned_converges_and_estimates_the_true_residual(oracle):
    rp, ci, va, A, b = _sys(10, 7)
    amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8)
    x, it, hist, conv = oracle.krylov("GMRES", rp, ci, va, b, amg=amg, tol=1e-9, max_iters=60, restart=20)
    assert conv and it < 30
    # the fixed AMG cycle is linear, so the Givens estimate |s[i+1]| equals the true residual norm up to rounding
    assert np.isclose(np.linalg.norm(b - A @ x), hist[-1], rtol=1e-5)
    assert np.all(np.diff(hist) <= 1e-14)      # GMRES residual estimates never increase inside a restart cycle


def test_gmres_single_iteration_path(oracle):
    """max_iters == 1 takes solve_one_iteration (gmres_solver.cu:215-268): x += (s0/H00) M^-1 v0"""
    rp, ci, va, A, b = _sys(7, 9)
    d = A.diagonal()
    x, it, hist, conv = oracle.krylov("GMRES", rp, ci, va, b, jacobi_omega=0.9, tol=1e-12, max_iters=1, restart=10)
    assert it == 1 and not conv
    z = 0.9 * (b / np.linalg.norm(b)) / d
    w = A @ z
    # minimise ||b - A z y||: y = <w, b> / <w, w>
    y = (w @ b) / (w @ w)
    assert np.allclose(x, y * z, rtol=1e-10)
    assert np.isclose(hist[1], np.linalg.norm(b - A @ x), rtol=1e-8)


def test_nonzero_initial_guess_and_immediate_convergence(oracle):
    rp, ci, va, A, b = _sys(6, 11)
    xt = np.linalg.solve(A.toarray(), b)
    for kind in ("CG", "PCGF", "PBICGSTAB", "GMRES"):
        x, it, hist, conv = oracle.krylov(kind, rp, ci, va, b, x0=xt * (1 + 1e-3), tol=1e-6, max_iters=80)
        assert conv and np.linalg.norm(b - A @ x) <= 1.01e-6 * hist[0]
        x, it, hist, conv = oracle.krylov(kind, rp, ci, va, np.zeros_like(b), tol=1e-6, max_iters=10)
        assert it == 0 and conv and not x.any()


def test_diagonal_symmetric_scaling_solves_the_original_system(oracle):
    """scaling = DIAGONAL_SYMMETRIC: the solver works on S A S, S b; the returned x is the solution of A x = b and the reported norms are
    those of the scaled system (src/solvers/solver.cu:667-675, 856-862)"""
    from tests.golden.cases_round2 import case_dict
    from tests.oracle_from_config import run_oracle
    (rp, ci, va), cfg = case_dict()["varpoisson12_pcg_agg_diagsym"]
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    b = np.random.default_rng(3).standard_normal(n)
    x, it, hist, conv, amg = run_oracle(oracle, cfg, rp, ci, va, b)
    assert conv
    s = 1.0 / np.sqrt(A.diagonal())
    assert np.isclose(hist[0], np.linalg.norm(s * b), rtol=1e-14)
    assert np.isclose(hist[-1], np.linalg.norm(s * (b - A @ x)), rtol=1e-6)
    assert np.linalg.norm(b - A @ x) <= 1e-7 * np.linalg.norm(b)
