"""Helpers shared by the opt-in GPU tests (components written after the round's GPU minutes were spent)."""
import os

import numpy as np
import pytest

UNVALIDATED = pytest.mark.validated_r2   # was an opt-in gate in round 1; every single-GPU suite ran on a B200 in round 2 (gpurun_out/r2/unvalidated.log)


def run_engine(amgx, cfgd, rp, ci, va, b, x0=None, mode="dDDI", block=1):
    """setup + solve through the C-ABI; returns (x, iterations, status, residual history, solver-less extras)"""
    n = rp.shape[0] - 1
    cfg = amgx.Config(cfgd)
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc, mode).upload(rp, ci, va)
    bv = amgx.Vector(rsc, mode).upload(b)
    xv = amgx.Vector(rsc, mode)
    if x0 is None:
        xv.set_zero(n * block)
    else:
        xv.upload(x0)
    slv = amgx.Solver(rsc, cfg, mode)
    try:
        slv.setup(A)
        slv.solve(bv, xv, zero_initial_guess=x0 is None)
        x = xv.download()
        try:
            hist = np.array(slv.residual_history()).ravel()
        except amgx.AMGXError:      # store_res_history off (the reference's own unit-test configurations): no history to read, as in the reference
            hist = np.zeros(0)
        return x, slv.iterations_number, slv.status, hist
    finally:
        for obj in (slv, xv, bv, A, rsc, cfg):
            obj.destroy()


def amg_agg_cfg(cycle="V", pre=1, post=1, omega=0.8, smoother="BLOCK_JACOBI", coarse="NOSOLVER", **extra):
    d = {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": cycle, "max_levels": 50,
         "presweeps": pre, "postsweeps": post, "coarse_solver": coarse, "max_iters": 1, "monitor_residual": 0,
         "smoother": {"scope": "sm", "solver": smoother, "relaxation_factor": omega, "monitor_residual": 0}}
    d.update(extra)
    return d


def outer_cfg(solver, precond, tol=1e-10, max_iters=60, **extra):
    s = {"scope": "main", "solver": solver, "max_iters": max_iters, "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI",
         "tolerance": tol, "norm": "L2", "preconditioner": precond}
    s.update(extra)
    return {"config_version": 2, "determinism_flag": 1, "solver": s}


NOPREC = {"scope": "nop", "solver": "NOSOLVER"}
JACOBI = {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}
