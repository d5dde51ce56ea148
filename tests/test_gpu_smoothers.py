"""GPU: MULTICOLOR_GS of the engine (csrc/gs.cu) against the CPU restatement.  Written after this round's GPU minutes were spent:
opt-in until validated on a device."""
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
from tests._gpu_util import UNVALIDATED, amg_agg_cfg, outer_cfg, run_engine

pytestmark = [pytest.mark.gpu, UNVALIDATED]


def _gs_smoother_cfg(sym, sweeps, w=0.9):
    return {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "MULTICOLOR_GS", "symmetric_GS": int(sym), "relaxation_factor": w, "matrix_coloring_scheme": "MIN_MAX",
        "max_iters": sweeps, "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": 1e-30, "norm": "L2"}}


@pytest.mark.parametrize("sym", [0, 1])
@pytest.mark.parametrize("mat", ["poisson", "banded"])
def test_gs_sweeps_bit_exact(amgx, oracle, sym, mat):
    """the kernel keeps the reference's lanes-per-row decomposition: every sweep must agree with the restatement bit for bit"""
    rp, ci, va = gallery.poisson7pt(13, 9, 7) if mat == "poisson" else sym_banded(5000, 60.0)
    n = rp.shape[0] - 1
    rng = np.random.default_rng(2)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    x, it, status, hist = run_engine(amgx, _gs_smoother_cfg(sym, 3), rp, ci, va, b, x0=x0)
    xo = x0
    for _ in range(3):
        xo = oracle.gs_sweep(rp, ci, va, b, xo, 0.9, symmetric=bool(sym))
    assert it == 3 and np.array_equal(x, xo)


@pytest.mark.parametrize("sym", [0, 1])
def test_amg_gs_matches_oracle(amgx, oracle, sym):
    rp, ci, va = gallery.poisson7pt(16, 14, 11)
    n = rp.shape[0] - 1
    b = np.ones(n)
    amg = amg_agg_cfg(smoother="MULTICOLOR_GS", omega=0.9)
    amg["smoother"].update(symmetric_GS=sym, matrix_coloring_scheme="MIN_MAX")
    amg["matrix_coloring_scheme"] = "MIN_MAX"
    outer = "PCG" if sym else "FGMRES"
    x, it, status, hist = run_engine(amgx, outer_cfg(outer, amg, tol=1e-9, max_iters=60, gmres_n_restart=20), rp, ci, va, b)
    oracle.set_uncolored_fraction(0.0)
    try:
        o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.9, smoother="MULTICOLOR_GS").set_symmetric_gs(bool(sym))
        if sym:
            xo, ito, histo, convo = oracle.pcg(rp, ci, va, b, amg=o, tol=1e-9, max_iters=60)
        else:
            xo, ito, histo, convo = oracle.fgmres(rp, ci, va, b, amg=o, tol=1e-9, max_iters=60, restart=20)
    finally:
        oracle.set_uncolored_fraction(0.15)
    assert convo and status == "success" and it == ito
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12


def _cheb_amg(precond, order, mode=2, coarsest=0, **extra):
    sm = {"scope": "cheb", "solver": "CHEBYSHEV", "max_iters": 1, "chebyshev_polynomial_order": order, "chebyshev_lambda_estimate_mode": mode,
          "monitor_residual": 0}
    if precond:
        sm["preconditioner"] = {"scope": "inner", "solver": precond, "max_iters": 1, "relaxation_factor": 0.9, "monitor_residual": 0}
    else:
        sm["preconditioner"] = {"scope": "inner", "solver": "NOSOLVER"}
    sm.update(extra)
    d = amg_agg_cfg(pre=0, post=1, error_scaling=3, coarsest_sweeps=coarsest)
    d["smoother"] = sm
    return d


@pytest.mark.parametrize("precond,order,mode", [(None, 2, 2), ("JACOBI_L1", 4, 2), ("BLOCK_JACOBI", 3, 3)])
def test_amg_chebyshev_matches_oracle(amgx, oracle, precond, order, mode):
    rp, ci, va = gallery.poisson7pt(15, 13, 10)
    n = rp.shape[0] - 1
    b = np.ones(n)
    # mode 3 (BLOCK_JACOBI(0.9)-preconditioned): the spectrum of 0.9 D^-1 A reaches 1.8, so the user bounds must cover it (0.95 stagnates, also in the oracle)
    amg = _cheb_amg(precond, order, mode, cheby_max_lambda=1.9, cheby_min_lambda=0.2)
    x, it, status, hist = run_engine(amgx, outer_cfg("PCG", amg, tol=1e-9, max_iters=80), rp, ci, va, b)
    oracle.set_chebyshev_precond(precond)
    try:
        o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=1, coarsest_sweeps=0, smoother="CHEBYSHEV")
    finally:
        oracle.set_chebyshev_precond(None)
    o.set_chebyshev(order=order, mode=mode, precond=precond, inner_omega=0.9, user_max=1.9, user_min=0.2).set_error_scaling(3)
    xo, ito, histo, convo = oracle.pcg(rp, ci, va, b, amg=o, tol=1e-9, max_iters=80)
    assert convo and status == "success" and it == ito
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12


def test_chebyshev_stale_x_on_the_coarsest_level_matches_oracle(amgx, oracle):
    """coarsest_sweeps > 0 with a Chebyshev smoother: the reference (and therefore engine and oracle) adds to whatever xc holds"""
    rp, ci, va = gallery.poisson7pt(10)
    n = rp.shape[0] - 1
    b = np.ones(n)
    amg = _cheb_amg(None, 2, coarsest=1)
    amg.update(scope="main", max_iters=6, monitor_residual=1, store_res_history=1, convergence="RELATIVE_INI", tolerance=1e-30, norm="L2")
    x, it, status, hist = run_engine(amgx, {"config_version": 2, "determinism_flag": 1, "solver": amg}, rp, ci, va, b)
    o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=1, coarsest_sweeps=1, smoother="CHEBYSHEV")
    o.set_chebyshev(order=2, mode=2).set_error_scaling(3)
    xo, ito, histo, convo = oracle.amg_solve(o, b, tol=1e-30, max_iters=6)
    assert it == ito == 6
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12


@pytest.mark.parametrize("order", [2, 4])
def test_amg_chebyshev_poly_matches_oracle(amgx, oracle, order):
    rp, ci, va = gallery.poisson7pt(14, 12, 11)
    n = rp.shape[0] - 1
    b = np.ones(n)
    amg = amg_agg_cfg(pre=0, post=3, coarsest_sweeps=0)
    amg["smoother"] = {"scope": "cp", "solver": "CHEBYSHEV_POLY", "chebyshev_polynomial_order": order, "max_iters": 1, "monitor_residual": 0}
    x, it, status, hist = run_engine(amgx, outer_cfg("PCG", amg, tol=1e-9, max_iters=80), rp, ci, va, b)
    o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, coarsest_sweeps=0, smoother="CHEBYSHEV_POLY").set_chebyshev(order=order)
    xo, ito, histo, convo = oracle.pcg(rp, ci, va, b, amg=o, tol=1e-9, max_iters=80)
    assert convo and status == "success" and it == ito
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12


def test_chebyshev_eigensolver_modes_fail_loudly(amgx):
    rp, ci, va = gallery.poisson7pt(5)
    with pytest.raises(Exception):
        run_engine(amgx, outer_cfg("PCG", _cheb_amg(None, 2, mode=0)), rp, ci, va, np.ones(125))


@pytest.mark.parametrize("mat", ["poisson", "banded"])
def test_parallel_greedy_colouring_bit_exact(amgx, oracle, mat):
    """integer work: the engine's colours on every level of a DILU hierarchy == the restatement's; solve history to 1e-12"""
    rp, ci, va = gallery.poisson7pt(15, 11, 9) if mat == "poisson" else sym_banded(6000, 50.0)
    n = rp.shape[0] - 1
    amg = amg_agg_cfg(pre=0, post=3, omega=0.75, smoother="MULTICOLOR_DILU", matrix_coloring_scheme="PARALLEL_GREEDY")
    amg["smoother"]["matrix_coloring_scheme"] = "PARALLEL_GREEDY"
    # the stand-alone AMG iteration diverges on the banded matrix (also in the oracle): there only the first sweeps are compared
    mi = 60 if mat == "poisson" else 4
    amg.update(scope="main", max_iters=mi, monitor_residual=1, store_res_history=1, convergence="RELATIVE_INI", tolerance=1e-8, norm="L2")
    cfg = amgx.Config({"config_version": 2, "determinism_flag": 1, "solver": amg})
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    b = amgx.Vector(rsc).upload(np.ones(n))
    x = amgx.Vector(rsc).set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    try:
        slv.setup(A)
        slv.solve(b, x, zero_initial_guess=True)
        hist = np.array(slv.residual_history()).ravel()
        cols = [slv.level_coloring(l) for l in range(slv.num_levels())]
        it, status = slv.iterations_number, slv.status
    finally:
        for obj in (slv, x, b, A, rsc, cfg):
            obj.destroy()
    oracle.set_coloring_scheme("PARALLEL_GREEDY")
    oracle.set_uncolored_fraction(0.0)
    try:
        o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.75, smoother="MULTICOLOR_DILU")
    finally:
        oracle.set_coloring_scheme("MIN_MAX")
        oracle.set_uncolored_fraction(0.15)
    assert len(cols) == o.num_levels()
    for l, (nc, colors) in enumerate(cols):
        nco, co, _ = oracle.amg_level_dilu(o, l)
        assert nc == nco and np.array_equal(colors, co), f"level {l} colouring"
    xo, ito, histo, convo = oracle.amg_solve(o, np.ones(n), tol=1e-8, max_iters=mi)
    assert it == ito and (status == "success") == bool(convo)
    if mat == "poisson":
        assert convo
    assert np.max(np.abs(hist - histo) / np.maximum(histo, histo[0])) < 1e-12
