"""ctypes binding of libamgx_b200.so -- the same calls an application makes through the
reference's C API (include/amgx_c.h), wrapped pyamgx-style (Config / Resources / Matrix / Vector /
Solver objects).  Host side only: all computation happens inside the CUDA library.  There is no CPU
fallback: if the library cannot be loaded the import of this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libamgx_b200.so"

# modes (include/amgx_config.h:102-124)
MODE = {"hDDI": 8192, "hDFI": 8448, "hFFI": 8464, "dDDI": 8193, "dDFI": 8449, "dFFI": 8465}
RC_NAMES = ["OK", "BAD_PARAMETERS", "UNKNOWN", "NOT_SUPPORTED_TARGET", "NOT_SUPPORTED_BLOCKSIZE", "CUDA_FAILURE",
            "THRUST_FAILURE", "NO_MEMORY", "IO_ERROR", "BAD_MODE", "CORE", "PLUGIN", "BAD_CONFIGURATION",
            "NOT_IMPLEMENTED", "LICENSE_NOT_FOUND", "INTERNAL"]
SOLVE_STATUS = {0: "success", 1: "failed", 2: "diverged", 3: "not_converged"}


class AMGXError(RuntimeError):
    def __init__(self, rc: int, where: str):
        self.rc = rc
        name = RC_NAMES[rc] if 0 <= rc < len(RC_NAMES) else str(rc)
        super().__init__(f"{where} failed with AMGX_RC_{name} ({rc})")


class AMGXB200_comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("world_size", C.c_int), ("nccl_unique_id", C.c_ubyte * 128)]

    def __init__(self, rank: int, world_size: int, unique_id: bytes):
        super().__init__()
        assert len(unique_id) == 128
        self.rank, self.world_size = rank, world_size
        C.memmove(C.addressof(self) + 8, unique_id, 128)   # raw copy: the id contains NUL bytes


class PartitionPlan(C.Structure):
    _fields_ = [("n_owned", C.c_int), ("n_interior", C.c_int), ("n_halo", C.c_int), ("num_neighbors", C.c_int),
                ("neighbors", C.POINTER(C.c_int)), ("send_offsets", C.POINTER(C.c_int)), ("send_maps", C.POINTER(C.c_int)),
                ("halo_offsets", C.POINTER(C.c_int)), ("halo_global", C.POINTER(C.c_int64)),
                ("perm_old_to_new", C.POINTER(C.c_int)), ("local_cols", C.POINTER(C.c_int))]


_lib = None


def _preload_nccl():
    """The engine links libnccl.so.2 by soname.  A process that later imports torch needs torch's bundled NCCL (newer than the system
    one): whichever copy is mapped first serves both, so map the bundled one first when it exists (no torch import needed)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            so = Path(list(spec.submodule_search_locations)[0]) / "lib" / "libnccl.so.2"
            if so.exists():
                C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    except Exception:
        pass      # the system libnccl is used


def load_library(path: str | os.PathLike | None = None) -> C.CDLL:
    """Load the engine.  Raises OSError when the shared library is missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else _LIB_PATH
    if not p.exists():
        raise OSError(f"{p} not found: build it with `python -m amgx_b200.build` (or __graft_entry__.build())")
    _preload_nccl()
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def _declare(lib):
    vp, ip, i = C.c_void_p, C.POINTER(C.c_int), C.c_int
    sig = {
        "AMGX_initialize": [], "AMGX_finalize": [], "AMGX_initialize_plugins": [], "AMGX_finalize_plugins": [],
        "AMGX_get_api_version": [ip, ip],
        "AMGX_get_error_string": [i, C.c_char_p, i],
        "AMGX_register_print_callback": [vp],
        "AMGX_pin_memory": [vp, C.c_uint],
        "AMGX_unpin_memory": [vp],
        "AMGX_config_create": [C.POINTER(vp), C.c_char_p],
        "AMGX_config_add_parameters": [C.POINTER(vp), C.c_char_p],
        "AMGX_config_create_from_file": [C.POINTER(vp), C.c_char_p],
        "AMGX_config_create_from_file_and_string": [C.POINTER(vp), C.c_char_p, C.c_char_p],
        "AMGX_config_get_default_number_of_rings": [vp, ip],
        "AMGX_config_destroy": [vp],
        "AMGX_resources_create": [C.POINTER(vp), vp, vp, i, ip],
        "AMGX_resources_create_simple": [C.POINTER(vp), vp],
        "AMGX_resources_destroy": [vp],
        "AMGX_matrix_create": [C.POINTER(vp), vp, i],
        "AMGX_matrix_destroy": [vp],
        "AMGX_matrix_upload_all": [vp, i, i, i, i, vp, vp, vp, vp],
        "AMGX_matrix_replace_coefficients": [vp, i, i, vp, vp],
        "AMGX_matrix_get_size": [vp, ip, ip, ip],
        "AMGX_matrix_get_nnz": [vp, ip],
        "AMGX_matrix_download_all": [vp, vp, vp, vp, C.POINTER(vp)],
        "AMGX_matrix_vector_multiply": [vp, vp, vp],
        "AMGX_matrix_attach_coloring": [vp, vp, i, i],
        "AMGX_matrix_upload_distributed": [vp, i, i, i, i, i, vp, vp, vp, vp, vp],
        "AMGX_matrix_upload_all_global": [vp, i, i, i, i, i, vp, vp, vp, vp, i, i, vp],
        "AMGX_distribution_create": [C.POINTER(vp), vp],
        "AMGX_distribution_destroy": [vp],
        "AMGX_distribution_set_partition_data": [vp, i, vp],
        "AMGX_distribution_set_32bit_colindices": [vp, i],
        "AMGX_vector_create": [C.POINTER(vp), vp, i],
        "AMGX_vector_destroy": [vp],
        "AMGX_vector_upload": [vp, i, i, vp],
        "AMGX_vector_set_zero": [vp, i, i],
        "AMGX_vector_set_random": [vp, i],
        "AMGX_vector_download": [vp, vp],
        "AMGX_vector_get_size": [vp, ip, ip],
        "AMGX_vector_bind": [vp, vp],
        "AMGX_solver_create": [C.POINTER(vp), vp, i, vp],
        "AMGX_solver_destroy": [vp],
        "AMGX_solver_setup": [vp, vp],
        "AMGX_solver_resetup": [vp, vp],
        "AMGX_solver_solve": [vp, vp, vp],
        "AMGX_solver_solve_with_0_initial_guess": [vp, vp, vp],
        "AMGX_solver_get_iterations_number": [vp, ip],
        "AMGX_solver_get_iteration_residual": [vp, i, i, C.POINTER(C.c_double)],
        "AMGX_solver_get_status": [vp, ip],
        "AMGX_solver_calculate_residual_norm": [vp, vp, vp, vp, vp],
        "AMGX_read_system": [vp, vp, vp, C.c_char_p],
        "AMGX_write_system": [vp, vp, vp, C.c_char_p],
        "AMGX_generate_distributed_poisson_7pt": [vp, vp, vp, i, i, i, i, i, i, i, i],
        "AMGXB200_get_nccl_unique_id": [C.c_char_p],
        "AMGXB200_solver_get_num_levels": [vp, ip],
        "AMGXB200_solver_get_level_info": [vp, i, ip, ip, ip, ip],
        "AMGXB200_solver_get_level_matrix": [vp, i, vp, vp, vp],
        "AMGXB200_solver_get_level_aggregates": [vp, i, vp, vp, vp],
        "AMGXB200_solver_get_level_P": [vp, i, ip, vp, vp, vp],
        "AMGXB200_solver_get_level_R": [vp, i, ip, vp, vp, vp],
        "AMGXB200_solver_get_level_cf_map": [vp, i, vp],
        "AMGXB200_solver_get_level_smoother_data": [vp, i, vp],
        "AMGXB200_solver_get_level_coloring": [vp, i, ip, vp],
        "AMGXB200_solver_get_last_solve_stats": [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)],
        "AMGXB200_bench_kernel": [vp, i, i, i, i, C.POINTER(C.c_double)],
        "AMGXB200_matrix_get_kernel_plan": [vp] + [C.POINTER(C.c_int)] * 5,
        "AMGXB200_partition_plan_create": [C.POINTER(PartitionPlan), i, i, vp, i, i, vp, vp],
        "AMGXB200_config_check": [vp, i, C.c_char_p, i],
        "AMGXB200_partition_vector_to_contiguous": [i, i, vp, vp, vp],
        "AMGXB200_comm_maps_to_global_cols": [i, i, vp, C.c_int64, i, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    lib.AMGXB200_partition_plan_free.argtypes = [C.POINTER(PartitionPlan)]
    lib.AMGXB200_partition_plan_free.restype = None


def _ck(rc: int, where: str):
    if rc != 0:
        raise AMGXError(rc, where)


def _ptr(a):
    """void* of a numpy array / torch tensor (host or device) / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor, host or CUDA: uploads use cudaMemcpyDefault
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


_PRINT_CB_TYPE = C.CFUNCTYPE(None, C.c_char_p, C.c_int)
_print_cb_keepalive = None


def initialize():
    _ck(load_library().AMGX_initialize(), "AMGX_initialize")


def finalize():
    _ck(load_library().AMGX_finalize(), "AMGX_finalize")


def register_print_callback(fn):
    """fn(str) or None to silence the library."""
    global _print_cb_keepalive
    if fn is None:
        def fn(_s):
            return None
    cb = _PRINT_CB_TYPE(lambda msg, n: fn(msg[:n].decode(errors="replace") if msg else ""))
    _print_cb_keepalive = cb
    _ck(load_library().AMGX_register_print_callback(C.cast(cb, C.c_void_p)), "AMGX_register_print_callback")


def error_string(rc: int) -> str:
    buf = C.create_string_buffer(256)
    load_library().AMGX_get_error_string(rc, buf, 256)
    return buf.value.decode()


class Config:
    def __init__(self, options: str | dict | None = None, file: str | None = None):
        import json
        self.lib = load_library()
        self.h = C.c_void_p()
        if isinstance(options, dict):
            options = json.dumps(options)
        if file is not None and options is not None:
            _ck(self.lib.AMGX_config_create_from_file_and_string(C.byref(self.h), file.encode(), options.encode()), "AMGX_config_create_from_file_and_string")
        elif file is not None:
            _ck(self.lib.AMGX_config_create_from_file(C.byref(self.h), file.encode()), "AMGX_config_create_from_file")
        else:
            _ck(self.lib.AMGX_config_create(C.byref(self.h), (options or "").encode()), "AMGX_config_create")

    def add_parameters(self, options: str):
        _ck(self.lib.AMGX_config_add_parameters(C.byref(self.h), options.encode()), "AMGX_config_add_parameters")

    def default_number_of_rings(self) -> int:
        n = C.c_int()
        _ck(self.lib.AMGX_config_get_default_number_of_rings(self.h, C.byref(n)), "AMGX_config_get_default_number_of_rings")
        return n.value

    def destroy(self):
        if self.h:
            self.lib.AMGX_config_destroy(self.h)
            self.h = C.c_void_p()


def config_check(cfg: "Config", mode: str = "dDDI"):
    """(supported, message): does the engine provide every component the configuration names?  Pure host code, no GPU needed."""
    lib = load_library()
    buf = C.create_string_buffer(1024)
    rc = lib.AMGXB200_config_check(cfg.h, MODE[mode], buf, 1024)
    return rc == 0, buf.value.decode(errors="replace")


class Resources:
    def __init__(self, cfg: Config, device: int | None = None, comm: AMGXB200_comm | None = None):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._comm = comm
        if device is None and comm is None:
            _ck(self.lib.AMGX_resources_create_simple(C.byref(self.h), cfg.h), "AMGX_resources_create_simple")
        else:
            dev = (C.c_int * 1)(device or 0)
            cp = C.cast(C.pointer(comm), C.c_void_p) if comm is not None else None
            _ck(self.lib.AMGX_resources_create(C.byref(self.h), cfg.h, cp, 1, dev), "AMGX_resources_create")

    def destroy(self):
        if self.h:
            self.lib.AMGX_resources_destroy(self.h)
            self.h = C.c_void_p()


class Matrix:
    def __init__(self, rsc: Resources, mode: str = "dDDI"):
        self.lib = load_library()
        self.mode = mode
        self.h = C.c_void_p()
        _ck(self.lib.AMGX_matrix_create(C.byref(self.h), rsc.h, MODE[mode]), "AMGX_matrix_create")

    @property
    def mat_dtype(self):
        return np.float64 if self.mode[2] == "D" else np.float32

    @property
    def vec_dtype(self):
        return np.float64 if self.mode[1] == "D" else np.float32

    def upload(self, row_ptrs, col_indices, data, diag_data=None, block_dims=(1, 1), n=None, nnz=None):
        """Host numpy arrays or device torch tensors (int32 / mat dtype)."""
        if n is None:
            n = int(row_ptrs.shape[0]) - 1
        if nnz is None:
            nnz = int(col_indices.shape[0])
        self._keep = (row_ptrs, col_indices, data, diag_data)
        _ck(self.lib.AMGX_matrix_upload_all(self.h, n, nnz, block_dims[0], block_dims[1], _ptr(row_ptrs), _ptr(col_indices), _ptr(data), _ptr(diag_data)),
            "AMGX_matrix_upload_all")
        self._keep = None
        return self

    def upload_scipy(self, A):
        A = A.tocsr()
        return self.upload(np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
                           np.ascontiguousarray(A.data, dtype=self.mat_dtype))

    def replace_coefficients(self, data, diag_data=None):
        n, _, _ = self.get_size()
        _ck(self.lib.AMGX_matrix_replace_coefficients(self.h, n, self.get_nnz(), _ptr(data), _ptr(diag_data)), "AMGX_matrix_replace_coefficients")

    def get_size(self):
        n, bx, by = C.c_int(), C.c_int(), C.c_int()
        _ck(self.lib.AMGX_matrix_get_size(self.h, C.byref(n), C.byref(bx), C.byref(by)), "AMGX_matrix_get_size")
        return n.value, bx.value, by.value

    def get_nnz(self):
        z = C.c_int()
        _ck(self.lib.AMGX_matrix_get_nnz(self.h, C.byref(z)), "AMGX_matrix_get_nnz")
        return z.value

    def download(self):
        n, bx, by = self.get_size()
        nnz = self.get_nnz()
        rp = np.empty(n + 1, np.int32)
        ci = np.empty(nnz, np.int32)
        va = np.empty(nnz * bx * by, self.mat_dtype)
        dp = C.c_void_p()
        _ck(self.lib.AMGX_matrix_download_all(self.h, _ptr(rp), _ptr(ci), _ptr(va), C.byref(dp)), "AMGX_matrix_download_all")
        diag = None
        if dp.value:
            diag = np.ctypeslib.as_array(C.cast(dp, C.POINTER(C.c_double if self.mat_dtype == np.float64 else C.c_float)), shape=(n * bx * by,)).copy()
            C.CDLL(None).free(dp)
        return rp, ci, va, diag

    def multiply(self, x: "Vector", y: "Vector"):
        _ck(self.lib.AMGX_matrix_vector_multiply(self.h, x.h, y.h), "AMGX_matrix_vector_multiply")

    def attach_coloring(self, colors: np.ndarray, num_colors: int):
        colors = np.ascontiguousarray(colors, dtype=np.int32)
        _ck(self.lib.AMGX_matrix_attach_coloring(self.h, _ptr(colors), colors.shape[0], num_colors), "AMGX_matrix_attach_coloring")

    def generate_poisson7(self, rhs: "Vector", sol: "Vector", nx, ny, nz, px=1, py=1, pz=1, rings=1):
        _ck(self.lib.AMGX_generate_distributed_poisson_7pt(self.h, rhs.h if rhs else None, sol.h if sol else None, 1, rings, nx, ny, nz, px, py, pz),
            "AMGX_generate_distributed_poisson_7pt")

    def kernel_info(self) -> dict:
        v = [C.c_int() for _ in range(5)]
        _ck(self.lib.AMGXB200_matrix_get_kernel_plan(self.h, *[C.byref(t) for t in v]), "AMGXB200_matrix_get_kernel_plan")
        return dict(zip(("tile_rows", "coded_tiles", "pair_tiles", "row_pattern_tiles", "window"), (t.value for t in v)))

    def bench_kernel(self, kind: int, warmup=3, reps=20, flush_l2=False) -> float:
        ms = C.c_double()
        _ck(self.lib.AMGXB200_bench_kernel(self.h, kind, warmup, reps, int(flush_l2), C.byref(ms)), "AMGXB200_bench_kernel")
        return ms.value

    def destroy(self):
        if self.h:
            self.lib.AMGX_matrix_destroy(self.h)
            self.h = C.c_void_p()


class Vector:
    def __init__(self, rsc: Resources, mode: str = "dDDI"):
        self.lib = load_library()
        self.mode = mode
        self.h = C.c_void_p()
        _ck(self.lib.AMGX_vector_create(C.byref(self.h), rsc.h, MODE[mode]), "AMGX_vector_create")

    @property
    def dtype(self):
        return np.float64 if self.mode[1] == "D" else np.float32

    def upload(self, data, block_dim: int = 1, n: int | None = None):
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, dtype=self.dtype)
        if n is None:
            n = int(data.shape[0]) // block_dim
        _ck(self.lib.AMGX_vector_upload(self.h, n, block_dim, _ptr(data)), "AMGX_vector_upload")
        return self

    def set_zero(self, n: int, block_dim: int = 1):
        _ck(self.lib.AMGX_vector_set_zero(self.h, n, block_dim), "AMGX_vector_set_zero")
        return self

    def get_size(self):
        n, bd = C.c_int(), C.c_int()
        _ck(self.lib.AMGX_vector_get_size(self.h, C.byref(n), C.byref(bd)), "AMGX_vector_get_size")
        return n.value, bd.value

    def download(self, out=None):
        n, bd = self.get_size()
        if out is None:
            out = np.empty(n * bd, self.dtype)
        _ck(self.lib.AMGX_vector_download(self.h, _ptr(out)), "AMGX_vector_download")
        return out

    def bind(self, A: Matrix):
        _ck(self.lib.AMGX_vector_bind(self.h, A.h), "AMGX_vector_bind")

    def destroy(self):
        if self.h:
            self.lib.AMGX_vector_destroy(self.h)
            self.h = C.c_void_p()


class Solver:
    def __init__(self, rsc: Resources, cfg: Config, mode: str = "dDDI"):
        self.lib = load_library()
        self.mode = mode
        self.h = C.c_void_p()
        _ck(self.lib.AMGX_solver_create(C.byref(self.h), rsc.h, MODE[mode], cfg.h), "AMGX_solver_create")

    def setup(self, A: Matrix):
        _ck(self.lib.AMGX_solver_setup(self.h, A.h), "AMGX_solver_setup")

    def resetup(self, A: Matrix):
        _ck(self.lib.AMGX_solver_resetup(self.h, A.h), "AMGX_solver_resetup")

    def solve(self, b: Vector, x: Vector, zero_initial_guess: bool = False):
        f = self.lib.AMGX_solver_solve_with_0_initial_guess if zero_initial_guess else self.lib.AMGX_solver_solve
        _ck(f(self.h, b.h, x.h), "AMGX_solver_solve")

    @property
    def status(self) -> str:
        st = C.c_int()
        _ck(self.lib.AMGX_solver_get_status(self.h, C.byref(st)), "AMGX_solver_get_status")
        return SOLVE_STATUS[st.value]

    @property
    def iterations_number(self) -> int:
        n = C.c_int()
        _ck(self.lib.AMGX_solver_get_iterations_number(self.h, C.byref(n)), "AMGX_solver_get_iterations_number")
        return n.value

    def get_residual(self, it: int, idx: int = 0) -> float:
        r = C.c_double()
        _ck(self.lib.AMGX_solver_get_iteration_residual(self.h, it, idx, C.byref(r)), "AMGX_solver_get_iteration_residual")
        return r.value

    def residual_history(self):
        return [self.get_residual(i) for i in range(self.iterations_number + 1)]

    def calculate_residual_norm(self, A: Matrix, b: Vector, x: Vector, block=1):
        out = np.zeros(block, A.vec_dtype)
        _ck(self.lib.AMGX_solver_calculate_residual_norm(self.h, A.h, b.h, x.h, _ptr(out)), "AMGX_solver_calculate_residual_norm")
        return out

    # ---- hierarchy introspection (AMGXB200 extensions) ----
    def num_levels(self) -> int:
        n = C.c_int()
        _ck(self.lib.AMGXB200_solver_get_num_levels(self.h, C.byref(n)), "AMGXB200_solver_get_num_levels")
        return n.value

    def level_info(self, lvl: int):
        n, nnz, bd, nc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _ck(self.lib.AMGXB200_solver_get_level_info(self.h, lvl, C.byref(n), C.byref(nnz), C.byref(bd), C.byref(nc)), "AMGXB200_solver_get_level_info")
        return dict(n=n.value, nnz=nnz.value, block_dim=bd.value, n_coarse=nc.value)

    def level_matrix(self, lvl: int):
        info = self.level_info(lvl)
        dt = np.float64 if self.mode[2] == "D" else np.float32
        rp = np.empty(info["n"] + 1, np.int32)
        ci = np.empty(info["nnz"], np.int32)
        va = np.empty(info["nnz"] * info["block_dim"] ** 2, dt)
        _ck(self.lib.AMGXB200_solver_get_level_matrix(self.h, lvl, _ptr(rp), _ptr(ci), _ptr(va)), "AMGXB200_solver_get_level_matrix")
        return rp, ci, va

    def level_aggregates(self, lvl: int):
        info = self.level_info(lvl)
        agg = np.empty(info["n"], np.int32)
        rp = np.empty(info["n_coarse"] + 1, np.int32)
        rc = np.empty(info["n"], np.int32)
        _ck(self.lib.AMGXB200_solver_get_level_aggregates(self.h, lvl, _ptr(agg), _ptr(rp), _ptr(rc)), "AMGXB200_solver_get_level_aggregates")
        return agg, rp, rc

    def _level_op(self, fn, lvl: int, nrows: int):
        nnz = C.c_int()
        _ck(fn(self.h, lvl, C.byref(nnz), None, None, None), fn.__name__)
        dt = np.float64 if self.mode[2] == "D" else np.float32
        rp = np.empty(nrows + 1, np.int32)
        ci = np.empty(nnz.value, np.int32)
        va = np.empty(nnz.value, dt)
        _ck(fn(self.h, lvl, C.byref(nnz), _ptr(rp), _ptr(ci), _ptr(va)), fn.__name__)
        return rp, ci, va

    def level_P(self, lvl: int):
        return self._level_op(self.lib.AMGXB200_solver_get_level_P, lvl, self.level_info(lvl)["n"])

    def level_R(self, lvl: int):
        return self._level_op(self.lib.AMGXB200_solver_get_level_R, lvl, self.level_info(lvl)["n_coarse"])

    def level_cf_map(self, lvl: int):
        cf = np.empty(self.level_info(lvl)["n"], np.int32)
        _ck(self.lib.AMGXB200_solver_get_level_cf_map(self.h, lvl, _ptr(cf)), "AMGXB200_solver_get_level_cf_map")
        return cf

    def level_smoother_data(self, lvl: int, count: int | None = None):
        info = self.level_info(lvl)
        dt = np.float64 if self.mode[2] == "D" else np.float32
        out = np.empty(count if count is not None else info["n"] * info["block_dim"] ** 2, dt)
        _ck(self.lib.AMGXB200_solver_get_level_smoother_data(self.h, lvl, _ptr(out)), "AMGXB200_solver_get_level_smoother_data")
        return out

    def level_coloring(self, lvl: int):
        nc = C.c_int()
        colors = np.empty(self.level_info(lvl)["n"], np.int32)
        _ck(self.lib.AMGXB200_solver_get_level_coloring(self.h, lvl, C.byref(nc), _ptr(colors)), "AMGXB200_solver_get_level_coloring")
        return nc.value, colors

    def last_solve_stats(self):
        s, k = C.c_double(), C.c_longlong()
        _ck(self.lib.AMGXB200_solver_get_last_solve_stats(self.h, C.byref(s), C.byref(k)), "AMGXB200_solver_get_last_solve_stats")
        return s.value, k.value

    def destroy(self):
        if self.h:
            self.lib.AMGX_solver_destroy(self.h)
            self.h = C.c_void_p()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _ck(load_library().AMGXB200_get_nccl_unique_id(buf), "AMGXB200_get_nccl_unique_id")
    return buf.raw
