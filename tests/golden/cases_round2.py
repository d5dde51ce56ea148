"""Case list for the components written after round 1 ran out of GPU minutes (no reference golden yet).  One definition drives
three runs: the reference (make_golden.py -> tests/golden/r2_<name>.npz), the CPU oracle (tests/oracle_from_config.py) and the
engine (tests/test_golden_round2.py).  Configs are in the reference's JSON format."""
from __future__ import annotations

import numpy as np

from amgx_b200 import gallery


def _outer(solver, precond, tol=1e-9, max_iters=80, **extra):
    s = {"scope": "main", "solver": solver, "max_iters": max_iters, "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI",
         "tolerance": tol, "norm": "L2", "preconditioner": precond}
    s.update(extra)
    return {"config_version": 2, "determinism_flag": 1, "solver": s}


def _agg(smoother=None, **extra):
    d = {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50, "presweeps": 1,
         "postsweeps": 1, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 1, "monitor_residual": 0, "print_grid_stats": 1,
         "smoother": smoother or {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}}
    d.update(extra)
    return d


def _standalone(amg, tol=1e-8, max_iters=60):
    a = dict(amg)
    a.update(scope="main", max_iters=max_iters, monitor_residual=1, store_res_history=1, convergence="RELATIVE_INI", tolerance=tol, norm="L2")
    return {"config_version": 2, "determinism_flag": 1, "solver": a}


NOP = {"scope": "nop", "solver": "NOSOLVER"}
JAC = {"scope": "pj", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}
GS = lambda sym: {"scope": "gs", "solver": "MULTICOLOR_GS", "relaxation_factor": 0.9, "symmetric_GS": sym, "matrix_coloring_scheme": "MIN_MAX",
                  "monitor_residual": 0}
CHEB = lambda inner, order, mode=2, **kw: dict({"scope": "cheb", "solver": "CHEBYSHEV", "max_iters": 1, "monitor_residual": 0,
                                               "chebyshev_polynomial_order": order, "chebyshev_lambda_estimate_mode": mode,
                                               "preconditioner": ({"scope": "inner", "solver": inner, "max_iters": 1, "relaxation_factor": 0.9,
                                                                   "monitor_residual": 0} if inner else {"scope": "inner", "solver": "NOSOLVER"})}, **kw)
CHEBP = lambda order: {"scope": "cp", "solver": "CHEBYSHEV_POLY", "max_iters": 1, "monitor_residual": 0, "chebyshev_polynomial_order": order}


def sym_banded(n, sigma):
    import scipy.sparse as sp  # noqa: F401
    rp, ci, va = gallery.random_banded(n, sigma=sigma)
    A = gallery.to_scipy(rp, ci, va)
    A = (A + A.T).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def cases():
    """yields (name, (row_ptr, col_idx, values), config dict)"""
    P = lambda *g: gallery.poisson7pt(*g)
    # Krylov drivers
    for kind in ("CG", "PCGF", "PBICGSTAB", "GMRES"):
        extra = {"gmres_n_restart": 7} if kind == "GMRES" else {}
        yield f"poisson12_{kind.lower()}_noprec", P(12), _outer(kind, NOP, max_iters=150, **extra)
        if kind != "CG":
            yield f"poisson12_{kind.lower()}_jacobi", P(12), _outer(kind, JAC, max_iters=150, **extra)
            yield f"poisson14x11x9_{kind.lower()}_agg", P(14, 11, 9), _outer(kind, _agg(), **extra)
    yield "banded3000_pbicgstab_agg", gallery.random_banded(3000, sigma=40.0), _outer("PBICGSTAB", _agg())
    yield "banded3000_gmres_agg", gallery.random_banded(3000, sigma=40.0), _outer("GMRES", _agg(), gmres_n_restart=10)
    yield "poisson9_gmres_one_iteration", P(9), _outer("GMRES", JAC, max_iters=1)
    # cycles
    for cyc in ("W", "F"):
        yield f"poisson16_pcg_agg_{cyc}", P(16), _outer("PCG", _agg(cycle=cyc))
    yield "poisson15x12x10_pcgf_agg_CG", P(15, 12, 10), _outer("PCGF", _agg(cycle="CG"))
    yield "poisson15x12x10_pcgf_agg_CGF", P(15, 12, 10), _outer("PCGF", _agg(cycle="CGF"))
    yield "poisson12_fgmres_agg_CG3", P(12), _outer("FGMRES", _agg(cycle="CG", cycle_iters=3), gmres_n_restart=10)
    yield "poisson12_amg_classical_CG", P(12), _standalone(
        {"solver": "AMG", "algorithm": "CLASSICAL", "selector": "PMIS", "interpolator": "D2", "cycle": "CG", "max_levels": 50, "presweeps": 1,
         "postsweeps": 1, "coarse_solver": "NOSOLVER", "print_grid_stats": 1,
         "smoother": {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}})
    # error scaling
    for es, steps, reuse in ((2, 2, 0), (3, 2, 0), (3, 0, 0), (3, 2, 2)):
        yield f"poisson16x13x9_pcg_agg_es{es}_s{steps}_r{reuse}", P(16, 13, 9), _outer(
            "PCG", _agg(error_scaling=es, scaling_smoother_steps=steps, reuse_scale=reuse))
    yield "poisson12_amg_agg_es3_pre0", P(12), _standalone(_agg(presweeps=0, postsweeps=2, error_scaling=3))
    # smoothers
    for sym in (0, 1):
        yield f"poisson16x14x11_amg_gs_sym{sym}", P(16, 14, 11), _standalone(_agg(GS(sym), matrix_coloring_scheme="MIN_MAX"))
    yield "symbanded3000_fgmres_agg_gs", sym_banded(3000, 40.0), _outer("FGMRES", _agg(GS(0), matrix_coloring_scheme="MIN_MAX"), gmres_n_restart=20)
    yield "poisson15x13x10_pcg_agg_cheb2", P(15, 13, 10), _outer("PCG", _agg(CHEB(None, 2), presweeps=0, coarsest_sweeps=0, error_scaling=3))
    yield "poisson15x13x10_pcg_agg_cheb4_l1", P(15, 13, 10), _outer("PCG", _agg(CHEB("JACOBI_L1", 4), presweeps=0, coarsest_sweeps=0, error_scaling=3))
    yield "poisson15x13x10_pcg_agg_cheb3_user", P(15, 13, 10), _outer(
        "PCG", _agg(CHEB("BLOCK_JACOBI", 3, 3, cheby_max_lambda=1.85, cheby_min_lambda=0.2), presweeps=0, coarsest_sweeps=0, error_scaling=3))
    yield "poisson10_amg_agg_cheb2_coarsest1", P(10), _standalone(_agg(CHEB(None, 2), presweeps=0, coarsest_sweeps=1, error_scaling=3), tol=1e-30, max_iters=6)
    for order in (2, 4):
        yield f"poisson14x12x11_pcg_agg_chebpoly{order}", P(14, 12, 11), _outer("PCG", _agg(CHEBP(order), presweeps=0, postsweeps=3, coarsest_sweeps=0))
    # HMIS selector (the two shipped HMIS configs: FGMRES_CLASSICAL_AGGRESSIVE_HMIS.json, AMG_CLASSICAL_L1_AGGRESSIVE_HMIS.json)
    cla = lambda **kw: dict({"scope": "amg", "solver": "AMG", "algorithm": "CLASSICAL", "selector": "HMIS", "interpolator": "D2", "aggressive_levels": 1,
                             "interp_max_elements": 4, "max_row_sum": 0.9, "strength_threshold": 0.25, "cycle": "V", "max_levels": 50, "min_coarse_rows": 2,
                             "presweeps": 2, "postsweeps": 2, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 1, "monitor_residual": 0,
                             "print_grid_stats": 1,
                             "smoother": {"scope": "jl1", "solver": "JACOBI_L1", "relaxation_factor": 1, "monitor_residual": 0}}, **kw)
    yield "poisson14_fgmres_classical_hmis_aggr", P(14), _outer("FGMRES", cla(), tol=1e-10, max_iters=60, gmres_n_restart=20)
    yield "poisson16x12x9_fgmres_classical_hmis", P(16, 12, 9), _outer("FGMRES", cla(aggressive_levels=0), tol=1e-10, max_iters=60, gmres_n_restart=20)
    yield "banded3000_fgmres_classical_hmis", gallery.random_banded(3000, sigma=40.0), _outer("FGMRES", cla(aggressive_levels=0), tol=1e-10, max_iters=40,
                                                                                                gmres_n_restart=20)
    # D1, the default interpolator (the shipped CLASSICAL_* configs that do not name one, e.g. AMG_CLASSICAL_PMIS.json, FGMRES_CLASSICAL_PMIS.json)
    d1 = lambda **kw: dict(cla(selector="PMIS", interpolator="D1", aggressive_levels=0, interp_max_elements=-1), **kw)
    yield "poisson14_fgmres_classical_d1", P(14), _outer("FGMRES", d1(), tol=1e-10, max_iters=60, gmres_n_restart=20)
    yield "poisson16x12x9_fgmres_classical_d1_trunc4", P(16, 12, 9), _outer("FGMRES", d1(interp_max_elements=4), tol=1e-10, max_iters=60, gmres_n_restart=20)
    yield "banded3000_fgmres_classical_d1", gallery.random_banded(3000, sigma=40.0), _outer("FGMRES", d1(), tol=1e-10, max_iters=40, gmres_n_restart=20)
    yield "poisson14_fgmres_classical_d1_hmis_aggr", P(14), _outer("FGMRES", d1(selector="HMIS", aggressive_levels=1), tol=1e-10, max_iters=60,
                                                                  gmres_n_restart=20)
    # PARALLEL_GREEDY colouring: the shipped FGMRES_AGGREGATION.json / FGMRES_AGGREGATION_DILU.json preconditioner (DILU, 0 + 3 sweeps,
    # DENSE_LU coarse solver, min_coarse_rows 32).  The reference's in-place colouring kernel is not reproducible run to run
    # (csrc/coloring.cu), so these goldens are compared loosely: see REFERENCE_NONDETERMINISTIC in tests/test_golden_round2.py
    dilu_pg = {"scope": "dilu", "solver": "MULTICOLOR_DILU", "relaxation_factor": 0.75, "matrix_coloring_scheme": "PARALLEL_GREEDY", "monitor_residual": 0}
    fa = _agg(dilu_pg, presweeps=0, postsweeps=3, coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=128, min_coarse_rows=32,
              matrix_coloring_scheme="PARALLEL_GREEDY")
    yield "poisson16_fgmres_agg_dilu_pgreedy", P(16), _outer("FGMRES", fa, tol=1e-8, max_iters=60, gmres_n_restart=10)
    yield "symbanded3000_fgmres_agg_dilu_pgreedy", sym_banded(3000, 40.0), _outer("FGMRES", fa, tol=1e-8, max_iters=60, gmres_n_restart=10)
    yield "poisson14_amg_gs_pgreedy", P(14), _standalone(_agg(dict(GS(1), matrix_coloring_scheme="PARALLEL_GREEDY"), matrix_coloring_scheme="PARALLEL_GREEDY"))
    # scaling = DIAGONAL_SYMMETRIC on the main solver (V-cheby-smoother.json): variable-coefficient SPD matrix so that S is not a multiple of I
    def var_poisson(n):
        rp, ci, va = gallery.poisson7pt(n)
        N = rp.shape[0] - 1
        dd = 1.0 + 0.75 * np.sin(0.37 * np.arange(N)) ** 2
        rows = np.repeat(np.arange(N), np.diff(rp))
        return rp, ci, va * dd[rows] * dd[ci]
    yield "varpoisson12_amg_agg_diagsym", var_poisson(12), _standalone(_agg(scaling="DIAGONAL_SYMMETRIC"))
    yield "varpoisson12_pcg_agg_diagsym", var_poisson(12), _outer("PCG", _agg(), scaling="DIAGONAL_SYMMETRIC")
    # SIZE_4 selector; the last one is the shipped AMG_AGGRREGATION_CG.json (SIZE_4, CG cycle, JACOBI_L1, 0 + 2 sweeps)
    yield "poisson16_pcg_agg_size4", P(16), _outer("PCG", _agg(selector="SIZE_4"))
    yield "banded3000_fgmres_agg_size4", gallery.random_banded(3000, sigma=40.0), _outer("FGMRES", _agg(selector="SIZE_4"), gmres_n_restart=20)
    yield "poisson14x12x9_amg_agg_size4_cgcycle", P(14, 12, 9), _standalone(
        _agg({"scope": "jl1", "solver": "JACOBI_L1", "relaxation_factor": 0.9, "monitor_residual": 0}, selector="SIZE_4", cycle="CG", presweeps=0,
             postsweeps=2, coarsest_sweeps=2, min_coarse_rows=2), tol=1e-6, max_iters=60)
    # dense LU coarse solver
    for rows in (32, 128):
        yield f"poisson12_pcg_agg_denselu{rows}", P(12), _outer("PCG", _agg(coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=rows))


def case_dict():
    return {name: (mat, cfg) for name, mat, cfg in cases()}
