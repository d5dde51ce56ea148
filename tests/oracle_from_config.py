"""Run the CPU oracle from a solver configuration written in the reference's JSON format (the same dict the engine and the
reference's ref_dump are given), so that one case definition drives all three: reference golden, oracle, engine.

Supported surface = what the oracle restates: outer PCG / PCGF / CG / PBICGSTAB / GMRES / FGMRES / stand-alone AMG; preconditioner
NOSOLVER / BLOCK_JACOBI / AMG (AGGREGATION SIZE_2 or CLASSICAL PMIS); smoothers BLOCK_JACOBI, JACOBI_L1, MULTICOLOR_DILU,
MULTICOLOR_GS, CHEBYSHEV, CHEBYSHEV_POLY; cycles V W F CG CGF; error_scaling; DENSE_LU_SOLVER / NOSOLVER coarse solver."""
from __future__ import annotations

import numpy as np

# defaults of the reference's parameter registry (src/core.cu) for the keys read here
DEFAULTS = dict(max_levels=100, presweeps=1, postsweeps=1, coarsest_sweeps=2, finest_sweeps=-1, cycle="V", cycle_iters=2, error_scaling=0,
                scaling_smoother_steps=2, reuse_scale=0, coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=128, min_coarse_rows=2,
                relaxation_factor=0.9, symmetric_GS=0, chebyshev_polynomial_order=5, chebyshev_lambda_estimate_mode=0, cheby_max_lambda=1.0,
                cheby_min_lambda=0.125, strength_threshold=0.25, max_row_sum=1.1, interpolator="D1", aggressive_levels=0,
                aggressive_interpolator="MULTIPASS", interp_max_elements=-1, gmres_n_restart=20, max_iters=100, tolerance=1e-12, norm="L2",
                algorithm="CLASSICAL", selector="PMIS", determinism_flag=0, max_uncolored_percentage=0.15)


def _get(d, k):
    return d.get(k, DEFAULTS[k])


def _sub(d, key, default_name):
    """a nested solver entry: either {"solver": NAME, ...} or just the NAME string"""
    v = d.get(key, default_name)
    return v if isinstance(v, dict) else {"solver": v}


def build_amg(oracle, a, rp, ci, va, determinism=0):
    sm = _sub(a, "smoother", "BLOCK_JACOBI")
    sname = sm["solver"]
    omega = sm.get("relaxation_factor", a.get("relaxation_factor", DEFAULTS["relaxation_factor"]))
    coarse = _sub(a, "coarse_solver", DEFAULTS["coarse_solver"])["solver"]
    kw = dict(max_levels=_get(a, "max_levels"), min_coarse_rows=_get(a, "min_coarse_rows"), presweeps=_get(a, "presweeps"),
              postsweeps=_get(a, "postsweeps"), coarsest_sweeps=_get(a, "coarsest_sweeps"), finest_sweeps=_get(a, "finest_sweeps"),
              smoother=sname, omega=omega, coarse_solver=coarse, dense_lu_num_rows=_get(a, "dense_lu_num_rows"))
    inner = None
    if sname == "CHEBYSHEV":
        pc = _sub(sm, "preconditioner", "NOSOLVER")
        inner = None if pc["solver"] == "NOSOLVER" else pc["solver"]
    uncol = 0.0 if determinism else sm.get("max_uncolored_percentage", a.get("max_uncolored_percentage", DEFAULTS["max_uncolored_percentage"]))
    scheme = sm.get("matrix_coloring_scheme", a.get("matrix_coloring_scheme", "MIN_MAX"))
    oracle.set_chebyshev_precond(inner)
    oracle.set_uncolored_fraction(uncol)
    oracle.set_coloring_scheme(scheme)
    try:
        if _get(a, "algorithm") == "AGGREGATION":
            assert a.get("selector", "SIZE_2") in ("SIZE_2", "SIZE_4")
            amg = oracle.AMG(rp, ci, va, selector=a.get("selector", "SIZE_2"), **kw)
        else:
            assert _get(a, "selector") in ("PMIS", "HMIS")
            amg = oracle.ClassicalAMG(rp, ci, va, selector=_get(a, "selector"), strength_threshold=_get(a, "strength_threshold"), max_row_sum=_get(a, "max_row_sum"),
                                      interpolator=_get(a, "interpolator"), aggressive_levels=_get(a, "aggressive_levels"),
                                      aggressive_interpolator=_get(a, "aggressive_interpolator"), interp_max_elements=_get(a, "interp_max_elements"), **kw)
    finally:
        oracle.set_chebyshev_precond(None)
        oracle.set_uncolored_fraction(0.15)
        oracle.set_coloring_scheme("MIN_MAX")
    amg.set_cycle(_get(a, "cycle")).set_cycle_iters(_get(a, "cycle_iters"))
    if _get(a, "algorithm") == "AGGREGATION":
        amg.set_error_scaling(_get(a, "error_scaling"), _get(a, "scaling_smoother_steps"), _get(a, "reuse_scale"))
    if sname == "MULTICOLOR_GS":
        amg.set_symmetric_gs(bool(sm.get("symmetric_GS", a.get("symmetric_GS", 0))))
    if sname in ("CHEBYSHEV", "CHEBYSHEV_POLY"):
        pc = _sub(sm, "preconditioner", "NOSOLVER")
        amg.set_chebyshev(order=_get(sm, "chebyshev_polynomial_order"), mode=_get(sm, "chebyshev_lambda_estimate_mode"), precond=inner,
                          inner_omega=pc.get("relaxation_factor", DEFAULTS["relaxation_factor"]), user_max=_get(sm, "cheby_max_lambda"),
                          user_min=_get(sm, "cheby_min_lambda"))
    return amg


def run_oracle(oracle, cfg, rp, ci, va, b, x0=None):
    """returns (x, iterations, residual history, converged, amg-or-None)"""
    s = cfg["solver"]
    if s.get("scaling", "NONE") == "DIAGONAL_SYMMETRIC":
        # src/solvers/solver.cu:440-477, 667-675, 856-862 + src/scalers/diagonal_symmetric.cu: the solver is set up and run on S A S with
        # S = diag(1/sqrt(a_ii)), b <- S b, x0 <- S^-1 x0, and the solution is scaled back; norms are those of the scaled system
        import copy
        n = rp.shape[0] - 1
        rows = np.repeat(np.arange(n), np.diff(rp))
        d = np.zeros(n)
        d[rows[ci == rows]] = va[ci == rows]
        sc = 1.0 / np.sqrt(d)
        vs = va * (sc[rows] * sc[ci])
        c2 = copy.deepcopy(cfg)
        c2["solver"]["scaling"] = "NONE"
        x, it, hist, conv, amg = run_oracle(oracle, c2, rp, ci, vs, b * sc, None if x0 is None else x0 / sc)
        return x * sc, it, hist, conv, amg
    assert s.get("scaling", "NONE") == "NONE", s["scaling"]
    det = cfg.get("determinism_flag", 0)
    name = s["solver"]
    tol, mi, norm = _get(s, "tolerance"), _get(s, "max_iters"), _get(s, "norm")
    if name == "AMG":
        amg = build_amg(oracle, s, rp, ci, va, det)
        x, it, hist, conv = oracle.amg_solve(amg, b, x0=x0, tol=tol, max_iters=mi, norm=norm)
        return x, it, hist, conv, amg
    pc = _sub(s, "preconditioner", "NOSOLVER")
    kw, amg = {}, None
    if pc["solver"] == "AMG":
        amg = build_amg(oracle, pc, rp, ci, va, det)
        kw["amg"] = amg
    elif pc["solver"] == "BLOCK_JACOBI":
        kw["jacobi_omega"] = pc.get("relaxation_factor", DEFAULTS["relaxation_factor"])
    else:
        assert pc["solver"] == "NOSOLVER", pc["solver"]
    if name == "PCG":
        x, it, hist, conv = oracle.pcg(rp, ci, va, b, x0=x0, tol=tol, max_iters=mi, norm=norm, **kw)
    elif name == "FGMRES":
        x, it, hist, conv = oracle.fgmres(rp, ci, va, b, x0=x0, tol=tol, max_iters=mi, restart=_get(s, "gmres_n_restart"), krylov_dim=s.get("gmres_krylov_dim", 0), **kw)
    else:
        x, it, hist, conv = oracle.krylov(name, rp, ci, va, b, x0=x0, tol=tol, max_iters=mi, restart=_get(s, "gmres_n_restart"), norm=norm, **kw)
    return x, it, hist, conv, amg
