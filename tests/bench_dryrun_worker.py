"""Dry run of bench.py's control flow without a GPU (test infrastructure, used by tests/test_bench_host_logic.py).

bench.py itself is untouched: this worker replaces `amgx_b200.capi` by a stand-in whose objects return made-up numbers, turns the
few torch.cuda calls of bench.py into no-ops and lets torch.distributed run over gloo, then calls bench.main().  What it checks is the
host logic around the measurements: argument handling, the order in which the JSON line is assembled, the context objects
(other_workloads, strong_512) and the guard that prints the line when a context object hangs.  It says nothing about the engine.

  BENCH_DRYRUN_HANG=strong   rank 1 never returns from its first solve of the strong_512 problem (guard test)
"""
import os
import sys
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402


# ---- torch: no device ----
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.synchronize = lambda *_a, **_k: None
torch.cuda.empty_cache = lambda *_a, **_k: None
torch.Tensor.cuda = lambda self, *_a, **_k: self


def _on_cpu(fn):
    def wrapped(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


torch.zeros = _on_cpu(torch.zeros)
torch.tensor = _on_cpu(torch.tensor)
_init = dist.init_process_group
dist.init_process_group = lambda _backend, **_k: _init("gloo")


# ---- amgx_b200.capi: made-up numbers ----
stub = types.ModuleType("amgx_b200.capi")
STATE = {"matrices": 0}


class AMGXB200_comm:
    def __init__(self, rank, world, uid):
        assert len(uid) == 128
        self.rank, self.world = rank, world


class _Obj:
    def destroy(self):
        pass


class Config(_Obj):
    h = 1

    def __init__(self, options=None, file=None):
        assert options is not None or Path(file).exists()


class Resources(_Obj):
    def __init__(self, cfg, device=None, comm=None):
        self.comm = comm


class Matrix(_Obj):
    h = 2

    def __init__(self, rsc, mode="dDDI"):
        self.rsc, self.n, self.nnz, self.bd = rsc, 0, 0, 1
        self.index = 0

    def generate_poisson7(self, rhs, sol, nx, ny, nz, px=1, py=1, pz=1, rings=1):
        STATE["matrices"] += 1
        self.index = STATE["matrices"]
        self.n, self.nnz = nx * ny * nz, 7 * nx * ny * nz
        rhs.n = sol.n = self.n

    def upload(self, rp, ci, va, diag_data=None, block_dims=(1, 1), n=None, nnz=None):
        self.n, self.nnz, self.bd = len(rp) - 1, len(ci), block_dims[0]
        self.index = 0
        return self

    def get_size(self):
        return self.n, self.bd, self.bd

    def get_nnz(self):
        return self.nnz

    def kernel_info(self):
        return {"tile_rows": 256, "coded_tiles": 7, "pair_tiles": 7, "row_pattern_tiles": 7, "window": 0}

    def bench_kernel(self, kind, warmup=3, reps=20, flush_l2=False):
        return 0.2 + 0.05 * kind


class Vector(_Obj):
    def __init__(self, rsc, mode="dDDI"):
        self.n = 0

    def upload(self, data, block_dim=1, n=None):
        self.n = len(data)
        return self

    def set_zero(self, n, block_dim=1):
        self.n = n * block_dim
        return self

    def bind(self, A):
        pass

    def download(self, out=None):
        if out is None:
            return np.zeros(self.n)
        out[:] = 0
        return out


class Solver(_Obj):
    def __init__(self, rsc, cfg, mode="dDDI"):
        self.iterations_number, self.status, self.A = 0, "success", None

    def setup(self, A):
        self.A = A

    def solve(self, b, x, zero_initial_guess=False):
        if os.environ.get("BENCH_DRYRUN_HANG") == "strong" and self.A.index >= 2 and os.environ.get("RANK") == "1":
            time.sleep(10_000)
        if dist.is_initialized():
            dist.barrier()                  # a distributed solve is collective
        self.iterations_number = 71

    def last_solve_stats(self):
        return 0.2, 7810

    def residual_history(self):
        return [1.0, 1e-3, 5e-7]

    def num_levels(self):
        return 3

    def level_info(self, lvl):
        return {"n": self.A.n >> lvl, "nnz": self.A.nnz >> lvl}


class _Lib:
    def AMGX_pin_memory(self, ptr, nbytes):
        return 0

    def AMGX_unpin_memory(self, ptr):
        return 0

    def AMGX_distribution_create(self, dh, cfg):
        return 0

    def AMGX_distribution_set_partition_data(self, dh, kind, ptr):
        return 0

    def AMGX_distribution_destroy(self, dh):
        return 0

    def AMGX_matrix_upload_distributed(self, A, n_global, n, nnz, bx, by, rp, ci, va, diag, dh):
        assert n_global > n > 0 and nnz > 0 and bx == by == 4
        return 0


stub.AMGXB200_comm, stub.Config, stub.Resources, stub.Matrix, stub.Vector, stub.Solver = AMGXB200_comm, Config, Resources, Matrix, Vector, Solver
stub.initialize = stub.finalize = lambda: None
stub.register_print_callback = lambda fn: None
stub.nccl_unique_id = lambda: bytes(128)
stub.load_library = lambda: _Lib()
import amgx_b200                  # noqa: E402
sys.modules["amgx_b200.capi"] = stub
amgx_b200.capi = stub

import bench                      # noqa: E402

if __name__ == "__main__":
    bench.main()
