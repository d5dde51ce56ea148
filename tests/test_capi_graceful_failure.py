"""The reference's CAPIFailure unit test (src/tests/capi_graceful_failure.cu) through the C-ABI: misuse returns an error code, never a
crash.  The part that needs no device runs on the CPU; the rest is a GPU test, opt-in until validated on a device."""
import ctypes as C

import numpy as np
import pytest

from amgx_b200 import capi, gallery
from tests._gpu_util import UNVALIDATED


@pytest.fixture(scope="module")
def lib():
    capi.initialize()
    return capi.load_library()


def test_host_side_failures(lib, tmp_path):
    cfg = C.c_void_p()
    assert lib.AMGX_config_create_from_file(C.byref(cfg), b"nonexisting_file.cfg") != 0
    assert lib.AMGX_config_create_from_file_and_string(C.byref(cfg), b"nonexisting_file.cfg", b"determinim_flag=1") != 0
    assert lib.AMGX_config_create_from_file_and_string(C.byref(cfg), b"nonexisting_file.cfg", None) != 0
    assert lib.AMGX_config_create(C.byref(cfg), None) != 0
    assert lib.AMGX_config_create(C.byref(cfg), b"determinism_flag=1") == 0
    assert lib.AMGX_config_destroy(cfg) == 0
    assert lib.AMGX_config_create(C.byref(cfg), b"bad_STRING = = = -1") != 0
    assert lib.AMGX_config_create(C.byref(cfg), b"") == 0
    assert lib.AMGX_config_destroy(cfg) == 0
    assert lib.AMGX_config_destroy(None) != 0
    assert lib.AMGX_solver_destroy(None) != 0
    assert lib.AMGX_vector_destroy(None) != 0
    assert lib.AMGX_matrix_destroy(None) != 0
    st = C.c_int()
    assert lib.AMGX_solver_get_status(None, C.byref(st)) != 0
    v, d, t = C.c_char_p(), C.c_char_p(), C.c_char_p()
    assert lib.AMGX_get_build_info_strings(C.byref(v), C.byref(d), C.byref(t)) == 0 and v.value


@pytest.mark.gpu
@UNVALIDATED
def test_device_side_failures(lib, tmp_path):
    cfg, rcfg = C.c_void_p(), C.c_void_p()
    assert lib.AMGX_config_create(C.byref(cfg), b"determinism_flag=1") == 0
    assert lib.AMGX_config_create(C.byref(rcfg), b"") == 0
    rsc = C.c_void_p()
    dev = C.c_int(0)
    assert lib.AMGX_resources_create(C.byref(rsc), rcfg, None, 1, C.byref(dev)) == 0
    dDDI = capi.MODES["dDDI"] if hasattr(capi, "MODES") else 8193
    A, b, x = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.AMGX_matrix_create(C.byref(A), rsc, -1) != 0
    assert lib.AMGX_matrix_create(C.byref(A), rsc, dDDI) == 0
    assert lib.AMGX_vector_create(C.byref(b), rsc, -1) != 0
    assert lib.AMGX_vector_create(C.byref(b), rsc, dDDI) == 0
    assert lib.AMGX_vector_create(C.byref(x), rsc, dDDI) == 0
    # a MatrixMarket file to read back (27-point in the reference; any system does)
    rp, ci, va = gallery.poisson7pt(6)
    n, nnz = rp.shape[0] - 1, ci.shape[0]
    fn = tmp_path / "temp_matrix.mtx"
    with open(fn, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{n} {n} {nnz}\n")
        for i in range(n):
            for k in range(rp[i], rp[i + 1]):
                f.write(f"{i + 1} {ci[k] + 1} {float(va[k])!r}\n")
    assert lib.AMGX_read_system(A, b, x, b"nonexisting_file.mtx") != 0
    assert lib.AMGX_read_system(A, b, x, str(fn).encode()) == 0
    nr, bx, by = C.c_int(), C.c_int(), C.c_int()
    assert lib.AMGX_matrix_get_size(A, C.byref(nr), C.byref(bx), C.byref(by)) == 0
    assert (nr.value, bx.value, by.value) == (n, 1, 1)
    rp32, ci32 = rp.astype(np.int32), ci.astype(np.int32)
    assert lib.AMGX_matrix_upload_all(A, n, nnz, 1, 1, rp32.ctypes.data, ci32.ctypes.data, va.ctypes.data, None) == 0
    assert lib.AMGX_matrix_upload_all(A, -1, nnz, 1, 1, rp32.ctypes.data, ci32.ctypes.data, va.ctypes.data, None) != 0
    assert lib.AMGX_matrix_destroy(A) == 0
    assert lib.AMGX_matrix_create(C.byref(A), rsc, dDDI) == 0
    assert lib.AMGX_read_system(A, b, x, str(fn).encode()) == 0
    assert lib.AMGX_matrix_replace_coefficients(A, n, nnz, None, None) == 0            # NULLs should be ok
    assert lib.AMGX_matrix_replace_coefficients(A, n, nnz, va.ctypes.data, None) == 0
    assert lib.AMGX_write_system(A, b, None, None) != 0
    assert lib.AMGX_write_system(None, b, None, str(tmp_path / "out.mtx").encode()) != 0
    slv = C.c_void_p()
    assert lib.AMGX_solver_create(C.byref(slv), rsc, -1, cfg) != 0
    assert lib.AMGX_solver_create(C.byref(slv), rsc, dDDI, cfg) == 0
    it, st, res = C.c_int(), C.c_int(), C.c_double()
    assert lib.AMGX_solver_get_iterations_number(slv, C.byref(it)) == 0
    assert lib.AMGX_solver_get_iteration_residual(slv, 0, 0, C.byref(res)) != 0        # store_res_history was not requested
    assert lib.AMGX_solver_get_status(slv, C.byref(st)) == 0
    assert st.value == 1                                                               # AMGX_SOLVE_FAILED (include/amgx_c.h:74-80)
    assert lib.AMGX_solver_destroy(slv) == 0
    assert lib.AMGX_vector_set_zero(b, -1, 1) != 0
    assert lib.AMGX_vector_destroy(b) == 0
    assert lib.AMGX_vector_create(C.byref(b), rsc, dDDI) == 0
    assert lib.AMGX_vector_set_zero(b, 2 * n, 1) == 0
    n1, b1 = C.c_int(), C.c_int()
    assert lib.AMGX_vector_get_size(b, C.byref(n1), C.byref(b1)) == 0 and (n1.value, b1.value) == (2 * n, 1)
    ones = np.ones(n)
    assert lib.AMGX_vector_upload(b, n, 1, ones.ctypes.data) == 0
    for h, fn_ in ((b, lib.AMGX_vector_destroy), (x, lib.AMGX_vector_destroy), (A, lib.AMGX_matrix_destroy), (rsc, lib.AMGX_resources_destroy),
                   (cfg, lib.AMGX_config_destroy), (rcfg, lib.AMGX_config_destroy)):
        assert fn_(h) == 0
