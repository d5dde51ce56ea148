"""Synthetic matrices of BASELINE.json's configs.  Host numpy only."""
from __future__ import annotations

import numpy as np


def poisson7pt(nx: int, ny: int | None = None, nz: int | None = None, dtype=np.float64):
    """7-point Poisson, row = i + nx*j + nx*ny*k; per row: diagonal 6 first, then -1 for
    i-1, i+1, j-1, j+1, k-1, k+1 when inside (entry order of the reference's generator,
    src/distributed/distributed_manager.cu:86-260).  Returns (row_ptr, col_idx, values)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    r = np.arange(n, dtype=np.int64)
    i = r % nx
    j = (r // nx) % ny
    k = r // (nx * ny)
    masks = [np.ones(n, bool), i > 0, i < nx - 1, j > 0, j < ny - 1, k > 0, k < nz - 1]
    offs = [0, -1, 1, -nx, nx, -nx * ny, nx * ny]
    cnt = np.zeros(n, np.int64)
    for m in masks:
        cnt += m
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=rp[1:])
    nnz = int(rp[-1])
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, dtype)
    pos = rp[:-1].copy()
    for m, o in zip(masks, offs):
        idx = pos[m]
        col[idx] = (r[m] + o).astype(np.int32)
        val[idx] = 6.0 if o == 0 else -1.0
        pos[m] += 1
    return rp.astype(np.int32), col, val


def poisson7pt_slab(nx: int, ny: int, nz: int, k0: int, k1: int, dtype=np.float64):
    """Rows of poisson7pt(nx, ny, nz) whose z index lies in [k0, k1): local row_ptr, GLOBAL (int64) columns, values.
    Used by the multi-GPU drivers: every rank builds only its own slab."""
    n_loc = nx * ny * (k1 - k0)
    r = np.arange(nx * ny * k0, nx * ny * k1, dtype=np.int64)
    i = r % nx
    j = (r // nx) % ny
    k = r // (nx * ny)
    masks = [np.ones(n_loc, bool), i > 0, i < nx - 1, j > 0, j < ny - 1, k > 0, k < nz - 1]
    offs = [0, -1, 1, -nx, nx, -nx * ny, nx * ny]
    cnt = np.zeros(n_loc, np.int64)
    for m in masks:
        cnt += m
    rp = np.zeros(n_loc + 1, np.int64)
    np.cumsum(cnt, out=rp[1:])
    nnz = int(rp[-1])
    col = np.empty(nnz, np.int64)
    val = np.empty(nnz, dtype)
    pos = rp[:-1].copy()
    for m, o in zip(masks, offs):
        idx = pos[m]
        col[idx] = r[m] + o
        val[idx] = 6.0 if o == 0 else -1.0
        pos[m] += 1
    return rp.astype(np.int32), col, val


def block_elasticity_slab(nx: int, ny: int, nz: int, k0: int, k1: int, dtype=np.float64):
    """Rows [k0, k1) (z slabs) of block_elasticity(nx, ny, nz) with GLOBAL block columns."""
    rp, col, val = poisson7pt_slab(nx, ny, nz, k0, k1)
    S = np.array([[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]], dtype=np.float64)
    I4 = np.eye(4)
    off = -(I4 + 0.1 * S)
    dia = 6.6 * I4 + 0.6 * S
    vals = np.empty((val.shape[0], 4, 4), dtype)          # filled in the target type: no fp64 temporary of nnz x 16
    vals[:] = off.astype(dtype)
    vals[val > 0] = dia.astype(dtype)
    return rp, col, vals.reshape(-1)


def poisson7pt_sorted(nx: int, ny: int | None = None, nz: int | None = None, dtype=np.float64):
    """Same matrix with ascending column order inside each row (cusp gallery order)."""
    rp, col, val = poisson7pt(nx, ny, nz, dtype)
    n = rp.shape[0] - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    order = np.lexsort((col, rows))
    return rp, col[order], val[order]


def _unique_sorted(key: np.ndarray) -> np.ndarray:
    """np.unique(key) -- on the GPU when one is there (the 60 M keys of the 4 M-row bench matrix sort in a second instead of a minute;
    the result is the same sorted set either way).  Host generator code only: nothing of the engine is involved."""
    if key.shape[0] > 5_000_000:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.unique(torch.from_numpy(key).cuda()).cpu().numpy()
        except Exception:
            pass
    key = np.sort(key)                      # (np.unique on 60 M int64 keys takes minutes with numpy 2.3; sort + neighbour mask: seconds)
    if key.shape[0] == 0:
        return key
    first = np.empty(key.shape[0], dtype=bool)
    first[0] = True
    np.not_equal(key[1:], key[:-1], out=first[1:])
    return key[first]


def random_banded(n: int = 4_000_000, seed: int = 12345, lam: float = 12.0, sigma: float = 2000.0, dtype=np.float64):
    """'SuiteSparse-shaped' CSR: row length 3+Poisson(lam) clipped to [1,64], columns
    row+N(0,sigma) deduplicated & clipped, off-diagonal U(-1,0), diagonal 1.05*sum|off| (first in row)."""
    rng = np.random.default_rng(seed)
    lens = np.clip(3 + rng.poisson(lam, n), 1, 64).astype(np.int64)
    tot = int(lens.sum())
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    cols = np.clip(rows + np.rint(rng.normal(0.0, sigma, tot)).astype(np.int64), 0, n - 1)
    keep = cols != rows
    rows, cols = rows[keep], cols[keep]
    key = rows * n + cols
    key = _unique_sorted(key)
    rows, cols = key // n, key % n
    vals = -rng.random(rows.shape[0])
    cnt = np.bincount(rows, minlength=n)
    offsum = np.bincount(rows, weights=-vals, minlength=n)
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(cnt + 1, out=rp[1:])
    nnz = int(rp[-1])
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, dtype)
    col[rp[:-1]] = np.arange(n, dtype=np.int32)
    val[rp[:-1]] = 1.05 * offsum + 1e-3
    starts = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=starts[1:])
    within = np.arange(rows.shape[0], dtype=np.int64) - starts[rows]
    dst = rp[rows] + 1 + within
    col[dst] = cols.astype(np.int32)
    val[dst] = vals
    return rp.astype(np.int32), col, val


def block_elasticity(nx: int, ny: int, nz: int, dtype=np.float64):
    """Block-4x4 7-point stencil: off-diagonal block -(I + 0.1 S), diagonal block
    6.6 I + 0.6 S, S a fixed symmetric 4x4 pattern; row-major blocks; diagonal block first."""
    rp, col, val = poisson7pt(nx, ny, nz)
    S = np.array([[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]], dtype=np.float64)
    I4 = np.eye(4)
    off = -(I4 + 0.1 * S)
    dia = 6.6 * I4 + 0.6 * S
    vals = np.empty((val.shape[0], 4, 4), dtype)          # filled in the target type: no fp64 temporary of nnz x 16
    vals[:] = off.astype(dtype)
    vals[val > 0] = dia.astype(dtype)
    return rp, col, vals.reshape(-1)


def to_scipy(rp, col, val, n=None):
    import scipy.sparse as sp
    n = rp.shape[0] - 1 if n is None else n
    return sp.csr_matrix((val, col, rp), shape=(n, n))
