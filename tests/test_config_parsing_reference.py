"""CPU: the known answers of the reference's own parser test (src/tests/config_parsing.cu: ConfigStringParsing) through the C-ABI.
The reference test feeds one AMG_Config object; here every string goes to a fresh AMGX_config_create (the expectations that depend on
nothing but the string itself are kept as they are; the two that relied on accumulated state are restated, see the comments)."""
import ctypes as C
import os

import pytest

from amgx_b200 import capi

OK, BAD = "ok", "AMGX_RC_BAD_CONFIGURATION"

STRINGS = [
    ("", OK),                                                                                  # empty string
    ("max_levels=10", OK),
    ("    max_levels = 10,min_coarse_rows = 10 ; \n max_iters \t= 10\n;", OK),
    ("    max_levels = 10 \n max_iters \t= 10\n", BAD),                                        # new line is not a delimiter
    ("    max_levels =  ,min_coarse_rows = 10", BAD),                                          # value not specified
    ("    max_levels = 10 min_coarse_rows = 10", BAD),                                         # no delimiter
    ("    definitely_nonexisting_parameter = 10, min_coarse_rows = 10", BAD),                  # bad parameter
    (" config_version=2,    solver(fgmres = 10, min_coarse_rows = 10", BAD),                   # unbalanced character
    ("  config_version=2,  solver(fgmres) = 10, fgmres:preconditioner(jacobi=BLOCK_JACOBI, min_coarse_rows = 10", BAD),
    ("    config_version=2, undefined_scope:max_iters = 10, min_coarse_rows = 10", OK),         # undefined scope
    ("    max_iters = 10, , min_coarse_rows = 10", OK),                                         # empty parameter
    ("    max_iters = 10,           , min_coarse_rows = 10", OK),
    ("    config_version=2, solver(scope)=FGMRES, preconditioner(scope)=BLOCK_JACOBI, max_iters = 10, min_coarse_rows = 10", BAD),   # two solvers, one scope
    ("    config_version=2, solver(scope2)=FGMRES, scope2::max_iters=10", BAD),                 # double colon
    ("    max_iters&=15", BAD),                                                                # invalid symbol
    ("    max_iters==15", BAD),                                                                # double equal
    ("    max_iters(scope3)=15", BAD),                                                         # new scope on a non-solver parameter (and version 1)
    ("    config_version=2, max_iters(scope3)=15", BAD),                                       # the same under config_version 2
    ("config_version=1, solver(scope)=FGMRES, scope:max_iters=1", BAD),                        # scopes need config_version 2
    ("config_version=3, max_iters=1", BAD),                                                    # config_version must be 1 or 2
    ("config_version=2, amg:determinism_flag=1", BAD),                                         # default-scope-only parameter
]

FILES = [
    ("     #SOME VERY LONG COMMENTS WITH ILLEGAL CHARACTER &^!@$!@)^*$::( \n  solver=FGMRES\n", OK),   # comments
    ("    \t  solver=FGMRES \n max_iters=10\n", OK),                                                       # tab
    ("     #           \n  solver=FGMRES\n", OK),                                                          # empty comment line
    ("    \n \n  solver=FGMRES \n max_iters=10 \n \n max_levels=10\n", OK),                                # blank lines
    ("\n", OK),
]


@pytest.fixture(scope="module")
def lib():
    capi.initialize()
    return capi.load_library()


@pytest.mark.parametrize("text,expected", STRINGS)
def test_config_strings(lib, text, expected):
    h = C.c_void_p()
    rc = lib.AMGX_config_create(C.byref(h), text.encode())
    assert (rc == 0) == (expected == OK), (rc, text)
    if expected != OK:
        assert rc == 12, rc           # AMGX_RC_BAD_CONFIGURATION (include/amgx_c.h:51-69)
    else:
        lib.AMGX_config_destroy(h)


@pytest.mark.parametrize("content,expected", FILES)
def test_config_files(lib, tmp_path, content, expected):
    fn = tmp_path / "temp.dat"
    fn.write_text(content)
    h = C.c_void_p()
    rc = lib.AMGX_config_create_from_file(C.byref(h), str(fn).encode())
    assert (rc == 0) == (expected == OK), (rc, content)
    if rc == 0:
        lib.AMGX_config_destroy(h)


def test_config_file_and_string_and_scope_redefinition(lib, tmp_path):
    fn = tmp_path / "temp.dat"
    body = "  config_version=2,  \n \n  solver(solver_name)=FGMRES \n solver_name:max_iters=10 \n \n max_levels=10\n"
    fn.write_text(body)
    h = C.c_void_p()
    assert lib.AMGX_config_create_from_file_and_string(C.byref(h), str(fn).encode(), b"config_version=2, solver_name:preconditioner=BLOCK_JACOBI") == 0
    # the reference test's last case parses the same file a second time into the same object and fails with "new scope already
    # defined"; through the C API the object is new every time, so the failure is restated inside one string ...
    g = C.c_void_p()
    assert lib.AMGX_config_create(C.byref(g), b"config_version=2, solver(solver_name)=FGMRES, preconditioner(solver_name)=AMG") == 12
    # ... while AMGX_config_add_parameters may redefine a scope (it parses with allow_configuration_mod, src/amgx_c.cu:2510-2515)
    assert lib.AMGX_config_add_parameters(C.byref(h), b"config_version=2, solver(solver_name)=PCG") == 0
    lib.AMGX_config_destroy(h)
    # a missing file
    assert lib.AMGX_config_create_from_file(C.byref(g), os.fsencode(tmp_path / "does_not_exist.cfg")) != 0
