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


# ---------------------------------------------------------------------------------------------------------------------------------
# FGMRES with gmres_krylov_dim < gmres_n_restart: the reference's truncated variant (orc_fgmres_trunc; fgmres_solver.cu:17-211, 406-569).
# No reference golden exists (no shipped configuration sets gmres_krylov_dim; parity unpinned): the restatement is pinned to a second,
# differently organised numpy formulation (no rings: every vector kept, H a dictionary that outlives the restarts) and to identities.
# ---------------------------------------------------------------------------------------------------------------------------------
def _nonsym(n=8, seed=4):
    rp, ci, va = gallery.poisson7pt(n)
    va = va.copy()
    rng = np.random.default_rng(seed)
    off = va < 0
    va[off] *= 0.4 + 0.6 * rng.random(int(off.sum()))          # nonsymmetric, diagonally dominant
    N = rp.shape[0] - 1
    return rp, ci, va, gallery.to_scipy(rp, ci, va), rng.standard_normal(N)


def _fgmres_trunc_numpy(A, b, dinv, omega, tol, max_iters, R, kd):
    K = min(min(max_iters, R), kd)
    N = b.shape[0]
    x = np.zeros(N)
    prec = (lambda v: omega * v * dinv) if dinv is not None else (lambda v: v.copy())
    H = {}
    h = lambda i, j: H.get((i, j), 0.0)
    cs, sn, s, gam = {}, {}, {}, {}
    nrm0 = np.linalg.norm(b)
    hist = [nrm0]
    V, Z, resid = {}, {}, None
    for it in range(max_iters):
        m = it % R
        if m == 0:
            V = {0: b - A @ x}
            Z = {}
            beta = np.linalg.norm(V[0])
            if it == 0 and beta <= tol * nrm0:
                hist.append(beta)
                return x, it + 1, hist
            V[0] = V[0] * (1.0 / beta)
            s = {0: beta}
        Z[m] = prec(V[m])
        w = A @ Z[m]
        lo = max(m - K, 0)
        for i in range(lo, m + 1):
            H[(i, m)] = V[i] @ w
            w = w - H[(i, m)] * V[i]
        H[(m + 1, m)] = np.linalg.norm(w)
        V[m + 1] = w * (1.0 / H[(m + 1, m)])
        gam[m] = s[m]
        for k in range(m):
            t = cs[k] * h(k, m) + sn[k] * h(k + 1, m)
            H[(k + 1, m)] = -sn[k] * h(k, m) + cs[k] * h(k + 1, m)
            H[(k, m)] = t
        dx, dy = h(m, m), h(m + 1, m)
        if dy < 0.0:
            cs[m], sn[m] = 1.0, 0.0
        elif abs(dy) > abs(dx):
            t = dx / dy
            sn[m] = 1.0 / np.sqrt(1.0 + t * t)
            cs[m] = t * sn[m]
        else:
            t = dy / dx
            cs[m] = 1.0 / np.sqrt(1.0 + t * t)
            sn[m] = t * cs[m]
        H[(m, m)] = cs[m] * dx + sn[m] * dy
        H[(m + 1, m)] = 0.0
        s[m + 1] = -sn[m] * s[m]
        s[m] = cs[m] * s[m]
        p = Z[m].copy()
        for i in range(lo, m):
            p -= h(i, m) * Z[i]
        p *= 1.0 / h(m, m)
        Z[m] = p
        x = x + s[m] * p
        if m == 0:
            resid = (s[1] * cs[0]) * V[1] + (-s[1] * sn[0]) * V[0]
        else:
            resid = (s[m + 1] * cs[m]) * V[m + 1] + (-s[m + 1] * sn[m] / gam[m]) * resid
        hist.append(np.linalg.norm(resid))
        if hist[-1] <= tol * nrm0:
            return x, it + 1, hist
    return x, max_iters, hist


@pytest.mark.parametrize("precond", ["none", "jacobi"])
@pytest.mark.parametrize("R,kd", [(30, 3), (7, 2), (12, 5)])
def test_fgmres_truncated_matches_numpy_formulation(oracle, precond, R, kd):
    """several restart cycles (R = 7, kd = 2 restarts five times): the columns m > K of the later cycles meet the rows the previous cycle
    left in H, as in the reference"""
    rp, ci, va, A, b = _nonsym()
    dinv = 1.0 / A.diagonal() if precond == "jacobi" else None
    kw = {"jacobi_omega": 0.8} if precond == "jacobi" else {}
    x, it, hist, conv = oracle.fgmres(rp, ci, va, b, tol=1e-8, max_iters=40, restart=R, krylov_dim=kd, **kw)
    xs, its, hs = _fgmres_trunc_numpy(A, b, dinv, 0.8, 1e-8, 40, R, kd)
    assert it == its and len(hist) == len(hs)
    assert np.allclose(hist, hs, rtol=1e-9, atol=1e-13 * hist[0])
    assert np.allclose(x, xs, rtol=1e-8, atol=1e-11)


def test_fgmres_truncated_without_actual_truncation_equals_fgmres(oracle):
    """krylov_dim below min(max_iters, restart) selects the variant, but a solve that ends within krylov_dim + 1 iterations never drops a
    vector: same iterates as FGMRES (x updated every iteration instead of once, residual norm from the recursion instead of |s[m+1]|)"""
    rp, ci, va, A, b = _nonsym(7, 9)
    x1, it1, h1, c1 = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-9, max_iters=60, restart=40)
    assert c1 and it1 < 38
    x2, it2, h2, c2 = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-9, max_iters=60, restart=40, krylov_dim=38)
    assert c2 and it2 == it1
    assert np.allclose(h1, h2, rtol=1e-7, atol=1e-12 * h1[0])
    assert np.allclose(x1, x2, rtol=1e-7, atol=1e-10)


def test_fgmres_truncated_residual_recursion_and_convergence(oracle):
    """while nothing has been dropped yet (m <= krylov_dim) the recursively updated residual is b - A x_m; afterwards the reference's
    p-recurrence ignores the fill-in of the R factor above the band, so the recursion only approximates the true residual -- the restart
    (true residual at m = 0) is what keeps the method honest, and it still converges"""
    rp, ci, va, A, b = _nonsym(8, 11)
    for kd in (3, 4, 6):
        j = kd + 1
        x, it, hist, conv = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-30, max_iters=j, restart=50, krylov_dim=kd)
        assert it == j and not conv
        assert abs(np.linalg.norm(b - A @ x) - hist[-1]) <= 1e-11 * hist[0]
    x, it, hist, conv = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-30, max_iters=12, restart=50, krylov_dim=3)
    assert abs(np.linalg.norm(b - A @ x) - hist[-1]) <= 1e-2 * hist[0]          # approximate once vectors have been dropped
    x, it, hist, conv = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-8, max_iters=300, restart=10, krylov_dim=2)
    xf, itf, hf, cf = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.8, tol=1e-8, max_iters=300, restart=10)
    assert conv and cf and it >= itf
    assert np.linalg.norm(b - A @ x) <= 1e-6 * hist[0]
