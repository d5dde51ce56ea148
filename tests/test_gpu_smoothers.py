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
