"""GPU: AMGX_matrix_replace_coefficients + AMGX_solver_resetup with structure_reuse_levels (aggregation hierarchies) against the CPU
restatement.  Written after this round's GPU minutes were spent: opt-in until validated on a device."""
import numpy as np
import pytest
import scipy.sparse as sp

from amgx_b200 import gallery
from tests._gpu_util import UNVALIDATED, amg_agg_cfg, outer_cfg

pytestmark = [pytest.mark.gpu, UNVALIDATED]


@pytest.mark.parametrize("k", [0, 1, 2, -1])
def test_resetup_structure_reuse_matches_oracle(amgx, oracle, k):
    rp, ci, va = gallery.poisson7pt(12, 10, 9)
    n = rp.shape[0] - 1
    A = gallery.to_scipy(rp, ci, va)
    A.sort_indices()
    D = 1.0 + 3.0 * np.random.default_rng(9).random(n)
    B = (sp.diags(D) @ A @ sp.diags(D)).tocsr()
    B.sort_indices()
    rp, ci, va, vb = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy(), B.data.copy()
    b = np.ones(n)
    cfg = amgx.Config(outer_cfg("PCG", amg_agg_cfg(structure_reuse_levels=k), tol=1e-9, max_iters=80))
    rsc = amgx.Resources(cfg)
    M = amgx.Matrix(rsc).upload(rp, ci, va)
    bv = amgx.Vector(rsc).upload(b)
    xv = amgx.Vector(rsc).set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    try:
        slv.setup(M)
        slv.solve(bv, xv, zero_initial_guess=True)
        kw = dict(max_levels=50, presweeps=1, postsweeps=1, omega=0.8)
        first = oracle.AMG(rp, ci, va, **kw)
        xo, ito, histo, convo = oracle.pcg(rp, ci, va, b, amg=first, tol=1e-9, max_iters=80)
        assert slv.iterations_number == ito
        M.replace_coefficients(vb)
        slv.resetup(M)
        xv.set_zero(n)
        slv.solve(bv, xv, zero_initial_guess=True)
        re = oracle.AMG(rp, ci, vb, reuse_from=first, structure_reuse_levels=k, **kw)
        xo, ito, histo, convo = oracle.pcg(rp, ci, vb, b, amg=re, tol=1e-9, max_iters=80)
        assert slv.num_levels() == re.num_levels()
        for l in range(re.num_levels() - 1):
            agg, _, _ = slv.level_aggregates(l)
            assert np.array_equal(agg, re.level(l)["aggregates"]), f"aggregates differ on level {l}"
        hist = np.array(slv.residual_history()).ravel()
        assert convo and slv.status == "success" and slv.iterations_number == ito
        assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12
    finally:
        for obj in (slv, xv, bv, M, rsc, cfg):
            obj.destroy()


@pytest.mark.parametrize("mat", ["poisson", "banded"])
def test_size4_aggregates_bit_exact(amgx, oracle, mat):
    """integer work: the SIZE_4 aggregates of every level == the restatement's; residual history to 1e-12"""
    rp, ci, va = gallery.poisson7pt(17, 13, 11) if mat == "poisson" else gallery.random_banded(5000, sigma=45.0)
    n = rp.shape[0] - 1
    b = np.ones(n)
    cfg = amgx.Config(outer_cfg("FGMRES", amg_agg_cfg(selector="SIZE_4"), tol=1e-9, max_iters=80, gmres_n_restart=20))
    rsc = amgx.Resources(cfg)
    M = amgx.Matrix(rsc).upload(rp, ci, va)
    bv = amgx.Vector(rsc).upload(b)
    xv = amgx.Vector(rsc).set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    try:
        slv.setup(M)
        slv.solve(bv, xv, zero_initial_guess=True)
        o = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8, selector="SIZE_4")
        assert slv.num_levels() == o.num_levels()
        for l in range(o.num_levels() - 1):
            agg, _, _ = slv.level_aggregates(l)
            assert np.array_equal(agg, o.level(l)["aggregates"]), f"aggregates differ on level {l}"
        xo, ito, histo, convo = oracle.fgmres(rp, ci, va, b, amg=o, tol=1e-9, max_iters=80, restart=20)
        hist = np.array(slv.residual_history()).ravel()
        assert convo and slv.status == "success" and slv.iterations_number == ito
        assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12
    finally:
        for obj in (slv, xv, bv, M, rsc, cfg):
            obj.destroy()


@pytest.mark.parametrize("k", [0, 2, -1])
def test_classical_resetup_keeps_P_and_R(amgx, oracle, k):
    """classical hierarchy: reused levels keep P and R (pattern and values), A_c = R A_new P"""
    from tests.golden.make_golden import cfg_fgmres_classical
    rp, ci, va = gallery.poisson7pt(12, 10, 9)
    n = rp.shape[0] - 1
    A0 = gallery.to_scipy(rp, ci, va)
    A0.sort_indices()
    D = 1.0 + 3.0 * np.random.default_rng(9).random(n)
    B = (sp.diags(D) @ A0 @ sp.diags(D)).tocsr()
    B.sort_indices()
    rp, ci, va, vb = A0.indptr.astype(np.int32), A0.indices.astype(np.int32), A0.data.copy(), B.data.copy()
    cfgd = cfg_fgmres_classical(aggressive_levels=0, tol=1e-9, max_iters=80)
    a = cfgd["solver"]["preconditioner"]
    a["structure_reuse_levels"] = k
    kw = dict(selector="PMIS", max_levels=a["max_levels"], min_coarse_rows=a["min_coarse_rows"], presweeps=a["presweeps"], postsweeps=a["postsweeps"],
              coarsest_sweeps=a["coarsest_sweeps"], smoother=a["smoother"]["solver"], omega=a["smoother"]["relaxation_factor"],
              strength_threshold=a["strength_threshold"], max_row_sum=a["max_row_sum"], interpolator=a["interpolator"],
              aggressive_levels=a["aggressive_levels"], interp_max_elements=a["interp_max_elements"])
    first = oracle.ClassicalAMG(rp, ci, va, **kw)
    o = oracle.ClassicalAMG(rp, ci, vb, reuse_from=first, structure_reuse_levels=k, **kw)
    s = cfgd["solver"]
    xo, ito, histo, convo = oracle.fgmres(rp, ci, vb, np.ones(n), amg=o, tol=s["tolerance"], max_iters=s["max_iters"], restart=s["gmres_n_restart"])
    cfg = amgx.Config(cfgd)
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    b, x = amgx.Vector(rsc).upload(np.ones(n)), amgx.Vector(rsc)
    x.set_zero(n)
    slv = amgx.Solver(rsc, cfg)
    slv.setup(A)
    A.replace_coefficients(vb)
    slv.resetup(A)
    slv.solve(b, x, zero_initial_guess=True)
    assert slv.num_levels() == o.num_levels()
    for l in range(slv.num_levels() - 1):
        assert np.array_equal(slv.level_matrix(l)[2], o.level(l)["values"]), (k, l)
        if l + 1 < slv.num_levels():
            assert np.array_equal(slv.level_P(l)[2], o.level(l)["P_values"]), (k, l)
    hist = np.array(slv.residual_history()).ravel()
    assert slv.iterations_number == ito and convo
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12
    for ob in (slv, x, b, A, rsc, cfg):
        ob.destroy()


def test_amg_levels_reuse_reference_unit_test(amgx):
    """the reference's AmgLevelsReuse unit test (src/tests/amg_levels_reuse.cu) with its own configuration string: 27-point Poisson, set up,
    solve, replace the coefficients, set up again (and, beyond the reference test, resetup), solve from the previous iterate -- the final x
    must not depend on structure_reuse_levels (1e-8)"""
    from tests.test_oracle_edge_cases import poisson27
    A = poisson27(24, 24, 24)
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    n = rp.shape[0] - 1
    rng = np.random.default_rng(1)
    b0, x0 = 1.0 + rng.random(n), 1.0 + np.random.default_rng(2).random(n)
    base = ("config_version=2, solver(main_solver)=AMG, main_solver:algorithm=AGGREGATION, main_solver:coarseAgenerator=LOW_DEG,main_solver:coloring_level=1,"
            "main_solver:convergence=RELATIVE_MAX,main_solver:cycle=V,main_solver:matrix_coloring_scheme=MIN_MAX,main_solver:max_levels=21,main_solver:norm=L1,"
            "main_solver:postsweeps=3,main_solver:presweeps=0,main_solver:selector=SIZE_2,main_solver:smoother=BLOCK_JACOBI,main_solver:tolerance=0.1,")
    tail = (",main_solver:max_iters=2,main_solver:monitor_residual=1,determinism_flag=1,max_uncolored_percentage=0.,main_solver:store_res_history=1,"
            "main_solver:obtain_timings=1")
    for use_resetup in (False, True):
        ref = None
        for k in range(0, 10, 3):
            cfg = amgx.Config(base + f", main_solver:structure_reuse_levels={k}, " + tail)
            rsc = amgx.Resources(cfg)
            M = amgx.Matrix(rsc).upload(rp, ci, va)
            bv, xv = amgx.Vector(rsc).upload(b0), amgx.Vector(rsc).upload(x0)
            slv = amgx.Solver(rsc, cfg)
            slv.setup(M)
            slv.solve(bv, xv)
            x1 = xv.download()
            M.replace_coefficients(va)
            (slv.resetup if use_resetup else slv.setup)(M)
            xv.upload(x1)
            slv.solve(bv, xv)
            x2 = xv.download()
            if ref is None:
                ref = x2
            else:
                assert np.max(np.abs(x2 - ref)) <= 1e-8 * max(1.0, np.max(np.abs(ref))), (use_resetup, k)
            for ob in (slv, xv, bv, M, rsc, cfg):
                ob.destroy()
