"""GPU: parity of the CUDA engine (through the C-ABI) with the CPU oracle on the same seeded inputs,
and directly with the reference's golden vectors.  Integer arrays bit-exact; fp64 within the
north-star tolerance (residual history 1e-12 relative); kernels whose per-row order is the
reference's (thread-per-row FMA chain) are required to be BIT-exact against the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from amgx_b200 import gallery

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
CFG = Path(__file__).resolve().parents[1] / "amgx_b200" / "configs"


def canon(rp, ci, va, n):
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    A.sort_indices()
    return A


def sym_banded(n, seed=7, sigma=30.0):
    rp, ci, va = gallery.random_banded(n, seed=seed, sigma=sigma)
    A = gallery.to_scipy(rp, ci, va)
    off = A - sp.diags(A.diagonal())
    S = (off + off.T) * 0.5
    S = S.tocsr()
    d = np.asarray(abs(S).sum(axis=1)).ravel() * 1.05 + 1e-3
    M = (S + sp.diags(d)).tocsr()
    M.sort_indices()
    return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)


def ragged(n, seed=3):
    """rows of length 0..40, some empty, no guaranteed diagonal"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 41, n)
    lens[rng.integers(0, n, n // 10)] = 0
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = rng.integers(0, n, rp[-1]).astype(np.int32)
    va = rng.standard_normal(rp[-1])
    return rp, ci, va


class Engine:
    def __init__(self, amgx, cfg, mode="dDDI"):
        self.amgx = amgx
        self.cfg = amgx.Config(cfg) if isinstance(cfg, (str, dict)) else cfg
        self.rsc = amgx.Resources(self.cfg)
        self.mode = mode
        self.objs = []

    def matrix(self, rp, ci, va, diag=None):
        A = self.amgx.Matrix(self.rsc, self.mode).upload(rp, ci, va, diag)
        self.objs.append(A)
        return A

    def vector(self, data=None, n=None):
        v = self.amgx.Vector(self.rsc, self.mode)
        if data is not None:
            v.upload(data)
        else:
            v.set_zero(n)
        self.objs.append(v)
        return v

    def solver(self):
        s = self.amgx.Solver(self.rsc, self.cfg, self.mode)
        self.objs.append(s)
        return s

    def close(self):
        for o in reversed(self.objs):
            o.destroy()
        self.rsc.destroy()
        self.cfg.destroy()


@pytest.fixture
def engine(amgx):
    made = []

    def make(cfg="config_version=2, solver(main)=NOSOLVER", mode="dDDI"):
        e = Engine(amgx, cfg, mode)
        made.append(e)
        return e
    yield make
    for e in made:
        e.close()


MATS = {
    "poisson20": lambda: gallery.poisson7pt(20),
    "poisson33x17x9": lambda: gallery.poisson7pt(33, 17, 9),
    "poisson_sorted12": lambda: gallery.poisson7pt_sorted(12),
    "banded20000": lambda: gallery.random_banded(20000, sigma=300.0),
    "ragged5000": lambda: ragged(5000),
    "tiny3": lambda: gallery.poisson7pt(3, 1, 1),
    "one_row": lambda: (np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.5])),
}


@pytest.mark.parametrize("name", list(MATS))
def test_spmv_bit_exact(engine, oracle, name):
    rp, ci, va = MATS[name]()
    n = rp.shape[0] - 1
    x = np.random.default_rng(1).standard_normal(n)
    e = engine()
    A, xv, yv = e.matrix(rp, ci, va), e.vector(x), e.vector(n=n)
    A.multiply(xv, yv)
    y = yv.download()
    assert np.array_equal(y, oracle.spmv(rp, ci, va, x))


def test_spmv_long_rows_fallback(engine, oracle):
    """rows longer than a shared-memory stage use the warp-per-row kernel: summation order differs -> tolerance"""
    rng = np.random.default_rng(5)
    n = 300
    lens = np.full(n, 10)
    lens[::7] = 30000 // 7
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = rng.integers(0, n, rp[-1]).astype(np.int32)
    va = rng.standard_normal(rp[-1])
    x = rng.standard_normal(n)
    e = engine()
    A, xv, yv = e.matrix(rp, ci, va), e.vector(x), e.vector(n=n)
    A.multiply(xv, yv)
    ref = oracle.spmv(rp, ci, va, x)
    assert np.max(np.abs(yv.download() - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_external_diagonal_upload_roundtrip(engine, oracle):
    rp, ci, va = gallery.poisson7pt(6)
    A = gallery.to_scipy(rp, ci, va)
    off = (A - sp.diags(A.diagonal())).tocsr()
    off.eliminate_zeros()
    off.sort_indices()
    d = A.diagonal().copy()
    e = engine()
    M = e.matrix(off.indptr.astype(np.int32), off.indices.astype(np.int32), off.data, diag=d)
    assert M.get_nnz() == off.nnz
    rp2, ci2, va2, d2 = M.download()
    assert np.array_equal(rp2, off.indptr) and np.array_equal(ci2, off.indices) and np.array_equal(va2, off.data) and np.array_equal(d2, d)
    x = np.random.default_rng(2).standard_normal(A.shape[0])
    xv, yv = e.vector(x), e.vector(n=A.shape[0])
    M.multiply(xv, yv)
    assert np.allclose(yv.download(), A @ x, rtol=1e-14, atol=1e-14)


def cfg_agg(tol=1e-10, max_iters=60, pre=0, post=3, omega=0.8, smoother="BLOCK_JACOBI", coarsest=2, norm="L2"):
    return {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "PCG", "max_iters": max_iters, "monitor_residual": 1, "store_res_history": 1,
        "convergence": "RELATIVE_INI", "tolerance": tol, "norm": norm,
        "preconditioner": {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V",
                           "max_levels": 50, "presweeps": pre, "postsweeps": post, "coarsest_sweeps": coarsest, "coarse_solver": "NOSOLVER",
                           "max_iters": 1, "monitor_residual": 0,
                           "smoother": {"scope": "jacobi", "solver": smoother, "relaxation_factor": omega, "monitor_residual": 0}}}}


def check_hierarchy(slv, amg, exact_values=True):
    assert slv.num_levels() == amg.num_levels()
    for l in range(amg.num_levels()):
        L = amg.level(l)
        info = slv.level_info(l)
        assert (info["n"], info["nnz"]) == (L["n"], L["nnz"]), f"level {l}"
        rp, ci, va = slv.level_matrix(l)
        A1, A2 = canon(rp, ci, va, L["n"]), canon(L["row_ptr"], L["col_idx"], L["values"], L["n"])
        assert np.array_equal(A1.indptr, A2.indptr) and np.array_equal(A1.indices, A2.indices)
        if exact_values:
            assert np.array_equal(A1.data, A2.data), f"level {l} values"
            assert np.array_equal(slv.level_smoother_data(l), L["d"]), f"level {l} smoother data"
        else:
            assert np.allclose(A1.data, A2.data, rtol=1e-13, atol=0)
        if l < amg.num_levels() - 1:
            agg, Rp, Rc = slv.level_aggregates(l)
            assert np.array_equal(agg, L["aggregates"]), f"level {l} aggregates"
            assert np.array_equal(Rp, L["R_row_offsets"]) and np.array_equal(Rc, L["R_column_indices"]), f"level {l} R"


@pytest.mark.parametrize("name,pre,post,omega,smoother", [
    ("poisson20", 0, 3, 0.8, "BLOCK_JACOBI"),
    ("poisson33x17x9", 0, 3, 0.8, "BLOCK_JACOBI"),
    ("poisson20", 2, 2, 0.7, "BLOCK_JACOBI"),
    ("poisson20", 1, 1, 1.0, "JACOBI_L1"),
    ("symbanded", 0, 3, 0.8, "BLOCK_JACOBI"),
])
def test_pcg_amg_vs_oracle(engine, oracle, name, pre, post, omega, smoother):
    rp, ci, va = sym_banded(8000) if name == "symbanded" else MATS[name]()
    n = rp.shape[0] - 1
    b = np.ones(n)
    e = engine(cfg_agg(pre=pre, post=post, omega=omega, smoother=smoother))
    A, bv, xv = e.matrix(rp, ci, va), e.vector(b), e.vector(n=n)
    slv = e.solver()
    slv.setup(A)
    slv.solve(bv, xv)
    amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=pre, postsweeps=post, smoother=smoother, omega=omega)
    check_hierarchy(slv, amg, exact_values=name.startswith("poisson"))
    xo, ito, histo, conv = oracle.pcg(rp, ci, va, b, amg=amg, tol=1e-10, max_iters=60)
    hist = np.array(slv.residual_history())
    assert slv.iterations_number == ito
    assert (slv.status == "success") == conv
    assert np.max(np.abs(hist - histo) / histo[0]) < 1e-12
    assert np.max(np.abs(xv.download() - xo)) <= 1e-10 * np.max(np.abs(xo))
    # independent check of the answer
    r = b - gallery.to_scipy(rp, ci, va) @ xv.download()
    assert abs(np.linalg.norm(r) - hist[-1]) <= 1e-8 * hist[0]


GOLDEN_AGG = ["poisson6_pcg_agg_jacobi", "poisson10_pcg_agg_jacobi", "poisson16_pcg_agg_jacobi", "poisson12x10x7_pcg_agg_jacobi",
              "poisson10_sorted_pcg_agg_jacobi", "poisson12_pcg_agg_jacobi_pre2"]


@pytest.mark.parametrize("name", GOLDEN_AGG)
def test_engine_vs_reference_golden(engine, name):
    """The CUDA engine against what the UNMODIFIED reference produced on a B200."""
    d = np.load(GOLD / f"{name}.npz")
    cfg = json.loads(str(d["config_json"]))
    cfg["solver"]["preconditioner"]["print_grid_stats"] = 0
    rp, ci, va, b = d["sys_row_ptr"], d["sys_col_idx"], d["sys_values"], d["sys_rhs"]
    n = rp.shape[0] - 1
    e = engine(cfg)
    A, bv, xv = e.matrix(rp, ci, va), e.vector(b), e.vector(n=n)
    slv = e.solver()
    slv.setup(A)
    slv.solve(bv, xv)
    nl = int(d["num_levels"][0])
    assert slv.num_levels() == nl
    for l in range(nl):
        info = d[f"L{l}.info"]
        li = slv.level_info(l)
        assert (li["n"], li["nnz"]) == (info[0], info[1])
        A1 = canon(*slv.level_matrix(l), li["n"])
        A2 = canon(d[f"L{l}.row_offsets"], d[f"L{l}.col_indices"], d[f"L{l}.values"][: info[1]], info[0])
        assert np.array_equal(A1.indptr, A2.indptr) and np.array_equal(A1.indices, A2.indices) and np.array_equal(A1.data, A2.data)
        assert np.array_equal(slv.level_smoother_data(l), d[f"L{l}.Dinv"])
        if l < nl - 1:
            agg, Rp, Rc = slv.level_aggregates(l)
            assert np.array_equal(agg, d[f"L{l}.aggregates"])
            assert np.array_equal(Rp, d[f"L{l}.R_row_offsets"]) and np.array_equal(Rc, d[f"L{l}.R_column_indices"])
    assert slv.iterations_number == int(d["iterations"][0])
    assert (slv.status == "success") == (int(d["status"][0]) == 0)
    hist = np.array(slv.residual_history())
    assert np.max(np.abs(hist - d["res_history"]) / d["res_history"][0]) < 1e-12
    assert np.max(np.abs(xv.download() - d["solution"])) <= 1e-10 * np.max(np.abs(d["solution"]))


def test_config1_example_matrix(engine):
    """BASELINE config 1: the shipped 12x12 example, PCG + BLOCK_JACOBI, vs the reference's own run."""
    d = np.load(GOLD / "example12_pcg_jacobi.npz")
    e = engine(json.loads(str(d["config_json"])))
    rp, ci, va, b = d["sys_row_ptr"], d["sys_col_idx"], d["sys_values"], d["sys_rhs"]
    A, bv, xv = e.matrix(rp, ci, va), e.vector(b), e.vector(n=12)
    slv = e.solver()
    slv.setup(A)
    slv.solve(bv, xv)
    assert slv.iterations_number == int(d["iterations"][0]) and slv.status == "not_converged"
    hist = np.array(slv.residual_history())
    assert np.allclose(hist[:6], d["res_history"][:6], rtol=1e-9)


def test_pcg_jacobi_golden(engine):
    d = np.load(GOLD / "poisson8_pcg_jacobi.npz")
    e = engine(json.loads(str(d["config_json"])))
    rp, ci, va, b = d["sys_row_ptr"], d["sys_col_idx"], d["sys_values"], d["sys_rhs"]
    A, bv, xv = e.matrix(rp, ci, va), e.vector(b), e.vector(n=rp.shape[0] - 1)
    slv = e.solver()
    slv.setup(A)
    slv.solve(bv, xv)
    assert slv.iterations_number == int(d["iterations"][0]) and slv.status == "success"
    hist = np.array(slv.residual_history())
    assert np.max(np.abs(hist - d["res_history"]) / d["res_history"][0]) < 1e-12


def test_solve_is_reproducible_and_reusable(engine):
    """same handle solved twice -> identical bits (deterministic reductions), and sol overwritten in place"""
    rp, ci, va = gallery.poisson7pt(24)
    n = rp.shape[0] - 1
    e = engine(cfg_agg(tol=1e-8))
    A, bv, xv = e.matrix(rp, ci, va), e.vector(np.ones(n)), e.vector(n=n)
    slv = e.solver()
    slv.setup(A)
    slv.solve(bv, xv)
    x1, h1 = xv.download(), slv.residual_history()
    xv.set_zero(n)
    slv.solve(bv, xv, zero_initial_guess=True)
    x2, h2 = xv.download(), slv.residual_history()
    assert np.array_equal(x1, x2) and h1 == h2
    assert np.array_equal(bv.download(), np.ones(n))           # rhs unchanged


def test_full_size_properties(engine):
    """128^3 (2.1 M rows): too big for the oracle in seconds -> size-independent properties:
    linearity of SpMV, residual norm of the returned solution recomputed independently, determinism."""
    nx = 128
    e = engine(cfg_agg(tol=1e-6, max_iters=100))
    A = e.amgx.Matrix(e.rsc)
    e.objs.append(A)
    bv, xv = e.amgx.Vector(e.rsc), e.amgx.Vector(e.rsc)
    e.objs += [bv, xv]
    A.generate_poisson7(bv, xv, nx, nx, nx)
    n = nx ** 3
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    uv, vv, wv, y1, y2, y3 = e.vector(u), e.vector(v), e.vector(2.0 * u - 3.0 * v), e.vector(n=n), e.vector(n=n), e.vector(n=n)
    A.multiply(uv, y1)
    A.multiply(vv, y2)
    A.multiply(wv, y3)
    lin = 2.0 * y1.download() - 3.0 * y2.download()
    assert np.max(np.abs(y3.download() - lin)) <= 1e-12 * np.max(np.abs(lin))
    # row sums of the Poisson operator: A*1 = number of missing neighbours
    ones = e.vector(np.ones(n))
    A.multiply(ones, y1)
    s = y1.download().reshape(nx, nx, nx)
    assert s[1:-1, 1:-1, 1:-1].max() == 0.0 and s.sum() == 6.0 * nx * nx
    slv = e.solver()
    slv.setup(A)
    xv.set_zero(n)
    slv.solve(bv, xv)
    assert slv.status == "success"
    hist = slv.residual_history()
    A.multiply(xv, y1)
    r = 1.0 - y1.download()
    assert abs(np.linalg.norm(r) - hist[-1]) <= 1e-9 * hist[0]
    assert hist[-1] <= 1e-6 * hist[0]
