"""GPU: the sliding-window CSR kernel (csrc/k_spmv_win.cu -- x of a CTA's tile range in a shared-memory ring, 16-bit column offsets) against
the CPU oracle.  The per-row FMA order is the oracle's, so SpMV must agree bit for bit; the window is a cache, so columns outside it (and
offsets that do not fit 16 bits) must give the same bits through the fallback path."""
import numpy as np
import pytest

from amgx_b200 import gallery
from tests._gpu_util import JACOBI, outer_cfg, run_engine
from tests.oracle_from_config import run_oracle

pytestmark = pytest.mark.gpu


def banded_with_outliers(n, sigma, seed=11, far=2000):
    """SuiteSparse-shaped banded matrix (the window plan accepts it: >= 148 * 8 tiles, ~16 entries per row, most columns within the window),
    then `far` off-diagonal entries moved anywhere in the matrix: beyond the window AND beyond 16-bit offsets"""
    rp, ci, va = gallery.random_banded(n, seed=seed, sigma=sigma)
    rng = np.random.default_rng(seed + 1)
    k = rng.integers(0, ci.shape[0], far)
    rows = np.searchsorted(rp, k, side="right") - 1
    k = k[ci[k] != rows]                                    # keep the diagonals
    ci = ci.copy()
    ci[k] = rng.integers(0, n, k.shape[0]).astype(np.int32)
    return rp, ci, va


@pytest.fixture(scope="module")
def system():
    return banded_with_outliers(400_003, 1500.0)            # odd row count: the last x entry cannot be bulk-copied


def test_window_plan_is_used(amgx, system):
    rp, ci, va = system
    cfg = amgx.Config("config_version=2, solver(main)=NOSOLVER")
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    try:
        info = A.kernel_info()
        assert info["window"] > 0 and info["tile_rows"] == 256, info
    finally:
        for o in (A, rsc, cfg):
            o.destroy()


def test_window_spmv_bit_exact(amgx, oracle, system):
    rp, ci, va = system
    n = rp.shape[0] - 1
    cfg = amgx.Config("config_version=2, solver(main)=NOSOLVER")
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    xv, yv = amgx.Vector(rsc), amgx.Vector(rsc)
    try:
        for seed in (1, 2):
            x = np.random.default_rng(seed).standard_normal(n)
            xv.upload(x)
            yv.set_zero(n)
            A.multiply(xv, yv)
            assert np.array_equal(yv.download(), oracle.spmv(rp, ci, va, x))
    finally:
        for o in (yv, xv, A, rsc, cfg):
            o.destroy()


def test_window_follows_replace_coefficients(amgx, oracle, system):
    """AMGX_matrix_replace_coefficients: the kernel's sliced-ELL copy of the values must follow the new coefficients"""
    rp, ci, va = system
    n = rp.shape[0] - 1
    cfg = amgx.Config("config_version=2, solver(main)=NOSOLVER")
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    xv, yv = amgx.Vector(rsc), amgx.Vector(rsc)
    try:
        x = np.random.default_rng(8).standard_normal(n)
        va2 = va * np.random.default_rng(9).uniform(0.5, 1.5, va.shape[0])
        A.replace_coefficients(va2)
        xv.upload(x)
        yv.set_zero(n)
        A.multiply(xv, yv)
        assert np.array_equal(yv.download(), oracle.spmv(rp, ci, va2, x))
    finally:
        for o in (yv, xv, A, rsc, cfg):
            o.destroy()


def test_window_jacobi_sweeps_bit_exact(amgx, oracle, system):
    """the fused Jacobi sweep of the window kernel (stand-alone BLOCK_JACOBI solver, 3 sweeps from a random x) against the oracle's sweeps"""
    rp, ci, va = system
    n = rp.shape[0] - 1
    rng = np.random.default_rng(4)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    cfg = {"config_version": 2, "solver": {"scope": "main", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 3, "monitor_residual": 0}}
    x, it, status, hist = run_engine(amgx, cfg, rp, ci, va, b, x0=x0)
    d = oracle.extract_diag(rp, ci, va)
    xo = x0
    for _ in range(3):
        xo = oracle.jacobi_sweep(rp, ci, va, d, b, xo, 0.8)
    assert np.max(np.abs(x - xo)) <= 1e-15 * np.max(np.abs(xo))      # same formula row by row (d^-1 by division, one FMA): last-bit agreement


def test_window_fgmres_jacobi_vs_oracle(amgx, oracle, system):
    """SpMV, residual + norm and the Krylov updates around the window kernel: FGMRES(6) + Jacobi history vs the oracle"""
    rp, ci, va = system
    n = rp.shape[0] - 1
    b = np.random.default_rng(3).standard_normal(n)
    cfg = outer_cfg("FGMRES", JACOBI, tol=1e-10, max_iters=12, gmres_n_restart=6)
    x, it, status, hist = run_engine(amgx, cfg, rp, ci, va, b)
    xo, ito, histo, conv, _ = run_oracle(oracle, cfg, rp, ci, va, b)
    assert it == ito
    m = min(len(hist), len(histo))
    assert m >= 2 and np.max(np.abs(hist[:m] - np.asarray(histo)[:m])) <= 1e-12 * hist[0]
    assert np.max(np.abs(x - xo)) <= 1e-11 * np.max(np.abs(xo))


def test_window_streaming_form_in_subprocess():
    """the opt-in streaming form (AMGXB_WINDOW_STREAM=1: per-warp chunk queues) through the same bit-exact checks; the switch is read once per
    process, hence the subprocess"""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    env = dict(os.environ, AMGXB_WINDOW_STREAM="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", str(root / "tests" / "test_gpu_window.py"), "-k",
                        "spmv_bit_exact or jacobi_sweeps or replace_coefficients"], capture_output=True, text=True, timeout=600, cwd=str(root), env=env)
    assert r.returncode == 0 and "3 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def banded_with_hub_rows(n, sigma, seed=21, hubs=900, hub_len=400):
    """the shape of the first aggregated level of the banded hierarchy: most rows short, ~1 % hub rows with hundreds of entries inside the band"""
    rp, ci, va = gallery.random_banded(n, seed=seed, sigma=sigma)
    rng = np.random.default_rng(seed + 1)
    hub = np.sort(rng.choice(np.arange(5000, n - 5000), hubs, replace=False))
    lens = np.diff(rp).astype(np.int64)
    new_lens = lens.copy()
    new_lens[hub] = hub_len
    rp2 = np.zeros(n + 1, np.int64)
    np.cumsum(new_lens, out=rp2[1:])
    ci2 = np.empty(rp2[-1], np.int32)
    va2 = np.empty(rp2[-1], np.float64)
    keep = np.ones(n, bool)
    keep[hub] = False
    # ordinary rows keep their entries
    src = np.repeat(rp[:-1][keep], lens[keep]) + (np.arange(lens[keep].sum()) - np.repeat(np.cumsum(lens[keep]) - lens[keep], lens[keep]))
    dst = np.repeat(rp2[:-1][keep], lens[keep]) + (np.arange(lens[keep].sum()) - np.repeat(np.cumsum(lens[keep]) - lens[keep], lens[keep]))
    ci2[dst] = ci[src]
    va2[dst] = va[src]
    for h in hub:                                            # hub rows: diagonal first, then distinct columns within +-4000
        cols = h + rng.choice(np.arange(-4000, 4001), hub_len - 1, replace=False)
        cols = cols[cols != h][:hub_len - 1]
        cols = np.concatenate([[h], cols, np.full(hub_len - 1 - cols.shape[0], h + 4001)])
        v = -rng.random(hub_len)
        v[0] = 1.05 * np.abs(v[1:]).sum() + 1e-3
        ci2[rp2[h]:rp2[h + 1]] = cols
        va2[rp2[h]:rp2[h + 1]] = v
    return rp2.astype(np.int32), ci2, va2


def test_window_long_rows_side_kernel(amgx, oracle):
    """hub rows leave the sliced-ELL copy and are summed by the warp-per-row side kernel (another summation order: tolerance); every other row
    stays bit-exact; the fused Jacobi sweep and the residual pick the parked dot products up"""
    rp, ci, va = banded_with_hub_rows(400_000, 1500.0)
    n = rp.shape[0] - 1
    lens = np.diff(rp)
    cfg = amgx.Config("config_version=2, solver(main)=NOSOLVER")
    rsc = amgx.Resources(cfg)
    A = amgx.Matrix(rsc).upload(rp, ci, va)
    xv, yv = amgx.Vector(rsc), amgx.Vector(rsc)
    try:
        assert A.kernel_info()["window"] > 0
        x = np.random.default_rng(5).standard_normal(n)
        xv.upload(x)
        yv.set_zero(n)
        A.multiply(xv, yv)
        y, ref = yv.download(), oracle.spmv(rp, ci, va, x)
        short = lens <= 64
        assert np.array_equal(y[short], ref[short])
        assert np.max(np.abs(y - ref)) <= 1e-13 * np.max(np.abs(ref))
    finally:
        for o in (yv, xv, A, rsc, cfg):
            o.destroy()
    rng = np.random.default_rng(6)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    jc = {"config_version": 2, "solver": {"scope": "main", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 2, "monitor_residual": 0}}
    xj, it, status, hist = run_engine(amgx, jc, rp, ci, va, b, x0=x0)
    d = oracle.extract_diag(rp, ci, va)
    xo = x0
    for _ in range(2):
        xo = oracle.jacobi_sweep(rp, ci, va, d, b, xo, 0.8)
    assert np.max(np.abs(xj - xo)) <= 1e-13 * np.max(np.abs(xo))
