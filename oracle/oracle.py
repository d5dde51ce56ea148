"""ctypes binding of the CPU oracle (oracle/amg_oracle.c).  TEST INFRASTRUCTURE: may be imported only
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(_HERE), "liboracle.so"], check=True)
    return _HERE / "liboracle.so"


def lib():
    global _lib
    if _lib is None:
        p = _HERE / "liboracle.so"
        if not p.exists() or p.stat().st_mtime < max(f.stat().st_mtime for f in list(_HERE.glob("*.c"))):
            build()
        _lib = C.CDLL(str(p))
        _lib.orc_dot.restype = C.c_double
        _lib.orc_nrm2.restype = C.c_double
        _lib.orc_nrm1.restype = C.c_double
        _lib.orc_nrmmax.restype = C.c_double
        _lib.orc_amg_setup.restype = C.c_void_p
        _lib.orc_amg_setup_classical.restype = C.c_void_p
        _lib.orc_cla_interpolate.restype = C.c_void_p
        _lib.orc_cla_galerkin.restype = C.c_void_p
    return _lib


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_num_threads(t: int):
    lib().orc_set_num_threads(C.c_int(t))


def num_threads() -> int:
    return lib().orc_num_threads()


def spmv(rp, ci, va, x):
    rp, ci, va, x = _i(rp), _i(ci), _d(va), _d(x)
    n = rp.shape[0] - 1
    y = np.empty(n)
    lib().orc_spmv(n, _p(rp), _p(ci), _p(va), _p(x), _p(y))
    return y


def residual(rp, ci, va, x, b):
    rp, ci, va, x, b = _i(rp), _i(ci), _d(va), _d(x), _d(b)
    n = rp.shape[0] - 1
    r = np.empty(n)
    lib().orc_residual(n, _p(rp), _p(ci), _p(va), _p(x), _p(b), _p(r))
    return r


def extract_diag(rp, ci, va):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    d = np.empty(n)
    lib().orc_extract_diag(n, _p(rp), _p(ci), _p(va), _p(d))
    return d


def l1_norms(rp, ci, va):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    d = np.empty(n)
    lib().orc_l1_norms(n, _p(rp), _p(ci), _p(va), _p(d))
    return d


def jacobi_sweep(rp, ci, va, d, b, x, omega):
    rp, ci, va, d, b, x = _i(rp), _i(ci), _d(va), _d(d), _d(b), _d(x)
    n = rp.shape[0] - 1
    out = np.empty(n)
    lib().orc_jacobi_sweep(n, _p(rp), _p(ci), _p(va), _p(d), _p(b), _p(x), _p(out), C.c_double(omega))
    return out


def jacobi_zero(d, b, omega):
    d, b = _d(d), _d(b)
    x = np.empty_like(b)
    lib().orc_jacobi_zero(b.shape[0], _p(d), _p(b), _p(x), C.c_double(omega))
    return x


def edge_weights(rp, ci, va, weight_formula=0):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    w = np.empty(ci.shape[0], np.float32)
    lib().orc_edge_weights(n, _p(rp), _p(ci), _p(va), weight_formula, _p(w))
    return w


def size2_aggregates(rp, ci, va, max_iterations=15, max_unassigned=0.05, merge_singletons=1, weight_formula=0):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    agg = np.empty(n, np.int32)
    nagg = lib().orc_size2_aggregates(n, _p(rp), _p(ci), _p(va), max_iterations, C.c_double(max_unassigned), merge_singletons, weight_formula, _p(agg))
    return agg, nagg


def size4_aggregates(rp, ci, va, max_iterations=15, max_unassigned=0.05, weight_formula=0):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    agg = np.empty(n, np.int32)
    nagg = lib().orc_size4_aggregates(n, _p(rp), _p(ci), _p(va), max_iterations, C.c_double(max_unassigned), weight_formula, _p(agg))
    return agg, nagg


def set_aggregation_selector(name):
    """selector of the NEXT aggregation setups: "SIZE_2" (default) or "SIZE_4"""
    lib().orc_set_aggregation_selector({"SIZE_2": 2, "SIZE_4": 4}[name])


def restriction(agg, nagg):
    agg = _i(agg)
    n = agg.shape[0]
    Rp = np.empty(nagg + 1, np.int32)
    Rc = np.empty(n, np.int32)
    lib().orc_restriction(n, nagg, _p(agg), _p(Rp), _p(Rc))
    return Rp, Rc


def galerkin(rp, ci, va, agg, nagg):
    rp, ci, va, agg = _i(rp), _i(ci), _d(va), _i(agg)
    n = rp.shape[0] - 1
    Rp, Rc = restriction(agg, nagg)
    rpc = np.empty(nagg + 1, np.int32)
    nnzc = lib().orc_galerkin_count(n, _p(rp), _p(ci), _p(agg), nagg, _p(Rp), _p(Rc), _p(rpc))
    cic = np.empty(nnzc, np.int32)
    vac = np.empty(nnzc)
    lib().orc_galerkin_fill(n, _p(rp), _p(ci), _p(va), _p(agg), nagg, _p(Rp), _p(Rc), _p(rpc), _p(cic), _p(vac))
    return rpc, cic, vac


SMOOTHERS = {"BLOCK_JACOBI": 0, "JACOBI_L1": 1, "MULTICOLOR_DILU": 2, "MULTICOLOR_GS": 3, "CHEBYSHEV": 4, "CHEBYSHEV_POLY": 5}


def set_chebyshev_precond(name):
    """which inner Jacobi the CHEBYSHEV smoother of the NEXT setup uses: None, "BLOCK_JACOBI" or "JACOBI_L1"""
    lib().orc_set_chebyshev_precond({None: 0, "BLOCK_JACOBI": 1, "JACOBI_L1": 2}[name])


class AMG:
    """Aggregation hierarchy + V-cycle exactly as the reference composes them (unfused)."""

    def __init__(self, rp, ci, va, max_levels=100, min_coarse_rows=2, coarsen_threshold=1.0, presweeps=1, postsweeps=1, coarsest_sweeps=2,
                 finest_sweeps=-1, smoother="BLOCK_JACOBI", omega=0.9, max_iterations=15, max_unassigned=0.05, merge_singletons=1, weight_formula=0,
                 coarse_solver="NOSOLVER", dense_lu_num_rows=128, reuse_from=None, structure_reuse_levels=0, selector="SIZE_2"):
        self.rp, self.ci, self.va = _i(rp), _i(ci), _d(va)
        set_aggregation_selector(selector)
        if reuse_from is not None:                       # AMGX_solver_resetup with structure_reuse_levels
            lib().orc_amg_reuse_structure(reuse_from.h, structure_reuse_levels)
        if coarse_solver == "DENSE_LU_SOLVER":
            min_coarse_rows = dense_lu_num_rows          # src/amg.cu:1154-1157
        self.n = self.rp.shape[0] - 1
        sm = SMOOTHERS[smoother]
        self.h = C.c_void_p(lib().orc_amg_setup(self.n, _p(self.rp), _p(self.ci), _p(self.va), max_levels, min_coarse_rows, C.c_double(coarsen_threshold),
                                                presweeps, postsweeps, coarsest_sweeps, finest_sweeps, sm, C.c_double(omega), max_iterations,
                                                C.c_double(max_unassigned), merge_singletons, weight_formula))
        set_aggregation_selector("SIZE_2")
        if coarse_solver == "DENSE_LU_SOLVER":
            lib().orc_amg_enable_dense_lu(self.h)

    def set_cycle(self, name: str):
        lib().orc_amg_set_cycle(self.h, {"V": 0, "W": 1, "F": 2, "CG": 3, "CGF": 4}[name])
        return self

    def set_chebyshev(self, order=5, mode=2, precond=None, inner_omega=0.9, user_max=1.0, user_min=0.125):
        """smoother CHEBYSHEV / CHEBYSHEV_POLY parameters; `precond` must match set_chebyshev_precond() at setup time"""
        lib().orc_amg_set_chebyshev(self.h, order, mode, {None: 0, "BLOCK_JACOBI": 1, "JACOBI_L1": 2}[precond], C.c_double(inner_omega),
                                    C.c_double(user_max), C.c_double(user_min))
        return self

    def level_lambda(self, l):
        a, b = C.c_double(), C.c_double()
        lib().orc_amg_level_lambda(self.h, l, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_symmetric_gs(self, sym: bool = True):
        lib().orc_amg_set_symmetric_gs(self.h, int(sym))
        return self

    def set_cycle_iters(self, iters: int):
        lib().orc_amg_set_cycle_iters(self.h, iters)
        return self

    def set_error_scaling(self, error_scaling: int, scaling_smoother_steps: int = 2, reuse_scale: int = 0):
        """aggregation levels only (the classical level ignores error_scaling in the reference too)"""
        lib().orc_amg_set_error_scaling(self.h, error_scaling, scaling_smoother_steps, reuse_scale)
        return self

    def num_levels(self):
        return lib().orc_amg_num_levels(self.h)

    def level(self, l):
        n, nnz, nagg = C.c_int(), C.c_int(), C.c_int()
        lib().orc_amg_level_sizes(self.h, l, C.byref(n), C.byref(nnz), C.byref(nagg))
        n, nnz, nagg = n.value, nnz.value, nagg.value
        rp = np.empty(n + 1, np.int32)
        ci = np.empty(nnz, np.int32)
        va = np.empty(nnz)
        d = np.empty(n)
        agg = np.empty(n, np.int32) if nagg else None
        Rp = np.empty(nagg + 1, np.int32) if nagg else None
        Rc = np.empty(n, np.int32) if nagg else None
        lib().orc_amg_level_arrays(self.h, l, _p(rp), _p(ci), _p(va), _p(agg) if nagg else None, _p(Rp) if nagg else None, _p(Rc) if nagg else None, _p(d))
        return dict(n=n, nnz=nnz, n_coarse=nagg, row_ptr=rp, col_idx=ci, values=va, aggregates=agg, R_row_offsets=Rp, R_column_indices=Rc, d=d)

    def vcycle(self, b, x=None):
        b = _d(b)
        zero = x is None
        x = np.zeros(self.n) if zero else _d(x).copy()
        lib().orc_amg_vcycle(self.h, _p(b), _p(x), int(zero))
        return x

    def __del__(self):
        try:
            if self.h:
                lib().orc_amg_free(self.h)
                self.h = None
        except Exception:
            pass


def pcg(rp, ci, va, b, amg: AMG | None = None, jacobi_omega: float | None = None, x0=None, tol=1e-6, max_iters=100, norm="L2"):
    rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
    n = rp.shape[0] - 1
    zero = x0 is None
    x = np.zeros(n) if zero else _d(x0).copy()
    hist = np.zeros(max_iters + 1)
    conv = C.c_int()
    precond = 1 if amg is not None else (2 if jacobi_omega is not None else 0)
    nt = {"L1": 0, "L2": 1, "LMAX": 2}[norm]
    it = lib().orc_pcg(n, _p(rp), _p(ci), _p(va), amg.h if amg is not None else None, precond, C.c_double(jacobi_omega or 0.0), _p(b), _p(x),
                       int(zero), C.c_double(tol), max_iters, nt, _p(hist), C.byref(conv))
    return x, it, hist[: it + 1].copy(), bool(conv.value)


def fgmres(rp, ci, va, b, amg: AMG | None = None, jacobi_omega: float | None = None, x0=None, tol=1e-6, max_iters=100, restart=20, krylov_dim=0):
    """krylov_dim (gmres_krylov_dim) smaller than min(max_iters, restart): the reference's truncated variant (orc_fgmres_trunc)"""
    rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
    n = rp.shape[0] - 1
    zero = x0 is None
    x = np.zeros(n) if zero else _d(x0).copy()
    hist = np.zeros(max_iters + 1)
    conv = C.c_int()
    precond = 1 if amg is not None else (2 if jacobi_omega is not None else 0)
    if 0 < krylov_dim < min(max_iters, restart):
        it = lib().orc_fgmres_trunc(n, _p(rp), _p(ci), _p(va), amg.h if amg is not None else None, precond, C.c_double(jacobi_omega or 0.0), _p(b), _p(x),
                                    int(zero), C.c_double(tol), max_iters, restart, int(krylov_dim), _p(hist), C.byref(conv))
        return x, it, hist[: it + 1].copy(), bool(conv.value)
    it = lib().orc_fgmres(n, _p(rp), _p(ci), _p(va), amg.h if amg is not None else None, precond, C.c_double(jacobi_omega or 0.0), _p(b), _p(x),
                          int(zero), C.c_double(tol), max_iters, restart, _p(hist), C.byref(conv))
    return x, it, hist[: it + 1].copy(), bool(conv.value)


def krylov(kind, rp, ci, va, b, amg: AMG | None = None, jacobi_omega: float | None = None, x0=None, tol=1e-6, max_iters=100, restart=20, norm="L2"):
    """CG / PCGF / PBICGSTAB / GMRES restatements (oracle/krylov_oracle.inc.c); same return convention as pcg()"""
    rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
    n = rp.shape[0] - 1
    zero = x0 is None
    x = np.zeros(n) if zero else _d(x0).copy()
    hist = np.zeros(max_iters + 1)
    conv = C.c_int()
    precond = 1 if amg is not None else (2 if jacobi_omega is not None else 0)
    nt = {"L1": 0, "L2": 1, "LMAX": 2}[norm]
    k = {"CG": 0, "PCGF": 1, "PBICGSTAB": 2, "GMRES": 3}[kind]
    it = lib().orc_krylov(k, n, _p(rp), _p(ci), _p(va), amg.h if amg is not None else None, precond, C.c_double(jacobi_omega or 0.0), _p(b), _p(x),
                          int(zero), C.c_double(tol), max_iters, restart, nt, _p(hist), C.byref(conv))
    return x, it, hist[: it + 1].copy(), bool(conv.value)


def set_uncolored_fraction(f: float):
    lib().orc_set_uncolored_fraction(C.c_double(f))


def amg_solve(amg: AMG, b, x0=None, tol=1e-6, max_iters=100, norm="L2"):
    b = _d(b)
    zero = x0 is None
    x = np.zeros(amg.n) if zero else _d(x0).copy()
    hist = np.zeros(max_iters + 1)
    conv = C.c_int()
    nt = {"L1": 0, "L2": 1, "LMAX": 2}[norm]
    it = lib().orc_amg_solve(amg.h, amg.n, _p(amg.rp), _p(amg.ci), _p(amg.va), _p(b), _p(x), int(zero), C.c_double(tol), max_iters, nt, _p(hist), C.byref(conv))
    return x, it, hist[: it + 1].copy(), bool(conv.value)


def amg_level_dilu(amg: AMG, l: int):
    n = amg.level(l)["n"]
    colors = np.empty(n, np.int32)
    einv = np.empty(n)
    nc = lib().orc_amg_level_dilu(amg.h, l, _p(colors), _p(einv))
    return nc, colors, einv


def gs_sweep(rp, ci, va, b, x, weight, symmetric=False, max_uncolored_fraction=0.0):
    """one MULTICOLOR_GS sweep (colours from MIN_MAX); returns the new x"""
    rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
    n = rp.shape[0] - 1
    nc, colors, srows, offs = color_min_max(rp, ci, max_uncolored_fraction)
    x = _d(x).copy()
    lib().orc_gs_sweep(n, _p(rp), _p(ci), _p(va), nc, _p(srows), _p(offs), _p(b), _p(x), C.c_double(weight), int(symmetric))
    return x


def set_coloring_scheme(name):
    """matrix_coloring_scheme of the multicolour smoothers of the NEXT setups: "MIN_MAX" (default) or "PARALLEL_GREEDY"""
    lib().orc_set_coloring_scheme({"MIN_MAX": 0, "PARALLEL_GREEDY": 1}[name])


def color_parallel_greedy(rp, ci, max_uncolored_fraction=0.0):
    rp, ci = _i(rp), _i(ci)
    n = rp.shape[0] - 1
    colors = np.empty(n, np.int32)
    nc = lib().orc_color_parallel_greedy(n, _p(rp), _p(ci), C.c_double(max_uncolored_fraction), _p(colors))
    sorted_rows = np.empty(n, np.int32)
    offsets = np.empty(nc + 1, np.int32)
    lib().orc_color_arrays(n, nc, _p(colors), _p(sorted_rows), _p(offsets))
    return nc, colors, sorted_rows, offsets


def color_min_max(rp, ci, max_uncolored_fraction=0.15):
    rp, ci = _i(rp), _i(ci)
    n = rp.shape[0] - 1
    colors = np.empty(n, np.int32)
    nc = lib().orc_color_min_max(n, _p(rp), _p(ci), C.c_double(max_uncolored_fraction), _p(colors))
    sorted_rows = np.empty(n, np.int32)
    offsets = np.empty(nc + 1, np.int32)
    lib().orc_color_arrays(n, nc, _p(colors), _p(sorted_rows), _p(offsets))
    return nc, colors, sorted_rows, offsets


def bspmv4(rp, ci, va, x):
    rp, ci, va, x = _i(rp), _i(ci), _d(va), _d(x)
    n = rp.shape[0] - 1
    y = np.empty(n * 4)
    lib().orc_bspmv4(n, _p(rp), _p(ci), _p(va), _p(x), _p(y))
    return y


def bjacobi4_dinv(rp, ci, va):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    d = np.empty(n * 16)
    lib().orc_bjacobi4_dinv(n, _p(rp), _p(ci), _p(va), _p(d))
    return d


def bjacobi4_sweep(rp, ci, va, dinv, b, x, weight):
    rp, ci, va, dinv, b, x = _i(rp), _i(ci), _d(va), _d(dinv), _d(b), _d(x)
    n = rp.shape[0] - 1
    out = np.empty(n * 4)
    lib().orc_bjacobi4_sweep(n, _p(rp), _p(ci), _p(va), _p(dinv), _p(b), _p(x), _p(out), C.c_double(weight))
    return out


def dilu4(rp, ci, va, b, x, weight, max_uncolored_fraction=0.15, sweeps=1):
    """colour (on the block graph), set up Einv and run `sweeps` DILU sweeps on a 4x4 block matrix; returns (x, Einv, colors)"""
    rp, ci, va, b = _i(rp), _i(ci), _d(va), _d(b)
    n = rp.shape[0] - 1
    nc, colors, srows, offs = color_min_max(rp, ci, max_uncolored_fraction)
    einv = np.empty(n * 16)
    lib().orc_dilu_setup_4x4(n, _p(rp), _p(ci), _p(va), nc, _p(colors), _p(srows), _p(offs), _p(einv))
    x = _d(x).copy()
    delta, Delta = np.zeros(n * 4), np.zeros(n * 4)
    for _ in range(sweeps):
        lib().orc_dilu_sweep_4x4(n, _p(rp), _p(ci), _p(va), nc, _p(colors), _p(srows), _p(offs), _p(einv), _p(b), _p(x), C.c_double(weight), _p(delta), _p(Delta))
    return x, einv, colors


# ---------------------------------------------------------------------------------------------------
# classical (Ruge-Stueben) AMG: strength, PMIS / aggressive PMIS, D2 / MULTIPASS, truncation, RAP
# ---------------------------------------------------------------------------------------------------
COARSE, FINE, STRONG_FINE, UNASSIGNED = -1, -2, -3, -4


def cla_strength(rp, ci, va, strength_threshold=0.25, max_row_sum=1.1):
    rp, ci, va = _i(rp), _i(ci), _d(va)
    n = rp.shape[0] - 1
    s_con = np.zeros(ci.shape[0], np.uint8)
    w = np.zeros(n, np.float32)
    lib().orc_cla_strength(n, _p(rp), _p(ci), _p(va), C.c_double(strength_threshold), C.c_double(max_row_sum), _p(s_con), _p(w))
    return s_con, w


def cla_pmis(rp, ci, s_con, weights, aggressive=False):
    """returns the C/F map with COARSE = -1 (not yet renumbered) -- weights is not modified"""
    rp, ci = _i(rp), _i(ci)
    n = rp.shape[0] - 1
    s_con = np.ascontiguousarray(s_con, np.uint8)
    w = np.ascontiguousarray(weights, np.float32).copy()
    cf = np.zeros(n, np.int32)
    if aggressive:
        lib().orc_cla_aggressive_pmis(n, _p(rp), _p(ci), _p(s_con), _p(w), _p(cf))
    else:
        lib().orc_cla_pmis(n, _p(rp), _p(ci), _p(s_con), _p(w), _p(cf), 0)
    return cf


def cla_rs(rp, ci, s_con, init=0):
    """Ruge-Stueben first pass (the host algorithm HMIS starts from); returns the C/F map"""
    rp, ci = _i(rp), _i(ci)
    n = rp.shape[0] - 1
    s_con = np.ascontiguousarray(s_con, np.uint8)
    cf = np.zeros(n, np.int32)
    lib().orc_cla_rs(n, _p(rp), _p(ci), _p(s_con), _p(cf), init)
    return cf


def cla_hmis(rp, ci, s_con, weights):
    rp, ci = _i(rp), _i(ci)
    n = rp.shape[0] - 1
    s_con = np.ascontiguousarray(s_con, np.uint8)
    w = np.ascontiguousarray(weights, np.float32).copy()
    cf = np.zeros(n, np.int32)
    lib().orc_cla_hmis(n, _p(rp), _p(ci), _p(s_con), _p(w), _p(cf))
    return cf


def set_classical_selector(name):
    """selector of the NEXT ClassicalAMG setups: "PMIS" (default) or "HMIS"""
    lib().orc_set_classical_selector({"PMIS": 0, "HMIS": 1}[name])


def cla_renumber(cf):
    cf = _i(cf).copy()
    nc = lib().orc_cla_renumber(cf.shape[0], _p(cf))
    return cf, nc


def _cla_take(h):
    n, nc, nnz = C.c_int(), C.c_int(), C.c_int()
    lib().orc_cla_matrix_sizes(h, C.byref(n), C.byref(nc), C.byref(nnz))
    rp = np.empty(n.value + 1, np.int32)
    ci = np.empty(nnz.value, np.int32)
    va = np.empty(nnz.value)
    lib().orc_cla_matrix_get(h, _p(rp), _p(ci), _p(va))
    return rp, ci, va, nc.value


def cla_interpolate(rp, ci, va, cf_renumbered, s_con, nc, interpolator="D2", max_elements=-1):
    rp, ci, va, cf = _i(rp), _i(ci), _d(va), _i(cf_renumbered)
    s_con = np.ascontiguousarray(s_con, np.uint8)
    n = rp.shape[0] - 1
    h = C.c_void_p(lib().orc_cla_interpolate(n, _p(rp), _p(ci), _p(va), _p(cf), _p(s_con), nc, {"D2": 0, "MULTIPASS": 1, "D1": 2}[interpolator], max_elements))
    out = _cla_take(h)
    lib().orc_cla_matrix_free(h)
    return out[:3]


class ClassicalAMG(AMG):
    """Classical hierarchy (PMIS / aggressive PMIS, D2 / MULTIPASS, truncation, R = P^T, RAP) + the same V-cycle."""

    def __init__(self, rp, ci, va, max_levels=100, min_coarse_rows=2, coarsen_threshold=1.0, presweeps=1, postsweeps=1, coarsest_sweeps=2,
                 finest_sweeps=-1, smoother="BLOCK_JACOBI", omega=0.9, strength_threshold=0.25, max_row_sum=1.1, interpolator="D1",
                 aggressive_levels=0, aggressive_interpolator="MULTIPASS", interp_max_elements=-1, coarse_solver="NOSOLVER", dense_lu_num_rows=128,
                 selector="PMIS", reuse_from=None, structure_reuse_levels=0):
        self.rp, self.ci, self.va = _i(rp), _i(ci), _d(va)
        if coarse_solver == "DENSE_LU_SOLVER":
            min_coarse_rows = dense_lu_num_rows
        self.n = self.rp.shape[0] - 1
        sm = SMOOTHERS[smoother]
        im = {"D2": 0, "MULTIPASS": 1, "D1": 2}
        set_classical_selector(selector)
        if reuse_from is not None:                       # AMGX_solver_resetup with structure_reuse_levels: P and R carried over
            lib().orc_amg_reuse_structure(reuse_from.h, structure_reuse_levels)
        self.h = C.c_void_p(lib().orc_amg_setup_classical(
            self.n, _p(self.rp), _p(self.ci), _p(self.va), max_levels, min_coarse_rows, C.c_double(coarsen_threshold), presweeps, postsweeps,
            coarsest_sweeps, finest_sweeps, sm, C.c_double(omega), C.c_double(strength_threshold), C.c_double(max_row_sum), im[interpolator],
            aggressive_levels, im[aggressive_interpolator], interp_max_elements))
        set_classical_selector("PMIS")
        if coarse_solver == "DENSE_LU_SOLVER":
            lib().orc_amg_enable_dense_lu(self.h)

    def level(self, l):
        d = super().level_plain(l) if hasattr(super(), "level_plain") else None
        n, nnz, nc = C.c_int(), C.c_int(), C.c_int()
        lib().orc_amg_level_sizes(self.h, l, C.byref(n), C.byref(nnz), C.byref(nc))
        n, nnz, nc = n.value, nnz.value, nc.value
        rp = np.empty(n + 1, np.int32)
        ci = np.empty(nnz, np.int32)
        va = np.empty(nnz)
        dd = np.empty(n)
        lib().orc_amg_level_arrays(self.h, l, _p(rp), _p(ci), _p(va), None, None, None, _p(dd))
        out = dict(n=n, nnz=nnz, n_coarse=nc, row_ptr=rp, col_idx=ci, values=va, d=dd)
        pnnz = C.c_int()
        if lib().orc_amg_level_classical(self.h, l, None, None, None, None, C.byref(pnnz)):
            cf = np.empty(n, np.int32)
            Pp = np.empty(n + 1, np.int32)
            Pc = np.empty(pnnz.value, np.int32)
            Pv = np.empty(pnnz.value)
            lib().orc_amg_level_classical(self.h, l, _p(cf), _p(Pp), _p(Pc), _p(Pv), None)
            out.update(cf_map=cf, P_row_offsets=Pp, P_col_indices=Pc, P_values=Pv)
        return out


def dense_lu_solve(A_dense, rhs):
    """LU with partial pivoting exactly as the engine's coarse solver orders it; returns (x, lu, ipiv)"""
    a = np.asfortranarray(np.array(A_dense, dtype=np.float64))
    n = a.shape[0]
    ipiv = np.zeros(n, np.int32)
    lib().orc_dense_lu_factor(n, _p(a), n, _p(ipiv))
    x = np.zeros(n)
    b = _d(rhs)
    lib().orc_dense_lu_solve(n, _p(a), n, _p(ipiv), _p(b), _p(x))
    return x, a, ipiv
