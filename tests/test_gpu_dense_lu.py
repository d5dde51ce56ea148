"""GPU: DENSE_LU_SOLVER (csrc/dense_lu.cu) against the CPU restatement.

First run on a B200 in round 2 (all green); part of the regular -m gpu suite since."""
import os

import numpy as np
import pytest

from amgx_b200 import gallery

pytestmark = [pytest.mark.gpu]


def cfg_pcg_agg_dense(num_rows=64, tol=1e-10):
    return {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "PCG", "max_iters": 60, "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI",
        "tolerance": tol, "norm": "L2",
        "preconditioner": {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
                           "presweeps": 0, "postsweeps": 3, "coarse_solver": "DENSE_LU_SOLVER", "dense_lu_num_rows": num_rows, "max_iters": 1,
                           "monitor_residual": 0, "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}}}}


@pytest.mark.parametrize("grid,num_rows", [((14, 14, 14), 64), ((20, 9, 13), 128), ((9, 9, 9), 2000)])
def test_pcg_amg_dense_lu_matches_oracle(amgx, oracle, grid, num_rows):
    rp, ci, va = gallery.poisson7pt(*grid)
    n = rp.shape[0] - 1
    cfgd = cfg_pcg_agg_dense(num_rows)
    cfg = amgx.Config(cfgd)
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    b = amgx.Vector(rsc).upload(np.ones(n))
    x = amgx.Vector(rsc).set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    slv.setup(A)
    slv.solve(b, x, zero_initial_guess=True)
    hist = np.array(slv.residual_history()).ravel()
    o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.8, coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=num_rows)
    assert slv.num_levels() == o.num_levels()
    xo, ito, histo, convo = oracle.pcg(rp, ci, va, np.ones(n), amg=o, tol=1e-10, max_iters=60)
    assert slv.iterations_number == ito and slv.status == "success" and convo
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12
    assert np.max(np.abs(x.download() - xo)) <= 1e-9 * np.max(np.abs(xo))
    for obj in (slv, x, b, A, rsc, cfg):
        obj.destroy()


def test_single_level_dense_lu_is_a_direct_solver(amgx):
    rp, ci, va = gallery.random_banded(300, sigma=20.0, seed=9)
    n = rp.shape[0] - 1
    cfgd = {"config_version": 2, "solver": {"scope": "main", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "max_levels": 1,
                                            "coarse_solver": "DENSE_LU_SOLVER", "dense_lu_num_rows": 128, "max_iters": 2, "monitor_residual": 1,
                                            "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": 1e-12, "norm": "L2",
                                            "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}}}
    cfg = amgx.Config(cfgd)
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    rhs = np.random.default_rng(4).standard_normal(n)
    b = amgx.Vector(rsc).upload(rhs)
    x = amgx.Vector(rsc).set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    slv.setup(A)
    slv.solve(b, x, zero_initial_guess=True)
    assert slv.iterations_number == 1 and slv.status == "success"
    M = gallery.to_scipy(rp, ci, va)
    assert np.linalg.norm(rhs - M @ x.download()) <= 1e-12 * np.linalg.norm(rhs)
    for obj in (slv, x, b, A, rsc, cfg):
        obj.destroy()
