"""CPU: AMGX_read_system_maps_one_ring / AMGX_read_system_global semantics through the resource-free hook
AMGXB200_read_system_partition, against the worked example the reference documents in examples/amgx_mpi_capi_agg.c:367-420
(12x12 matrix, three partitions [0 0 0 0 1 1 1 1 2 2 2 2]) -- a golden vector from the reference's own sources."""
import ctypes as C

import numpy as np
import pytest

from amgx_b200 import capi

RP = [0, 4, 8, 13, 21, 25, 32, 36, 41, 46, 50, 57, 61]
CI = [0, 1, 3, 8, 0, 1, 2, 3, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 8, 10, 2, 4, 5, 6, 2, 3, 4, 5, 6, 7, 10, 4, 5, 6, 7, 5, 6, 7, 9, 10, 0, 3, 8, 10, 11,
      7, 9, 10, 11, 3, 5, 7, 8, 9, 10, 11, 8, 9, 10, 11]
PV = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
# the outputs the reference documents for partitions 0, 1, 2
EXPECT = {
    0: dict(n=4, nnz=21, rp=[0, 4, 8, 13, 21], ci=[0, 1, 3, 6, 0, 1, 2, 3, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 7], nb=[1, 2],
            send=[[2, 3], [0, 3]], recv=[[4, 5], [6, 7]]),
    1: dict(n=4, nnz=20, rp=[0, 4, 11, 15, 20], ci=[4, 0, 1, 2, 4, 5, 0, 1, 2, 3, 7, 0, 1, 2, 3, 1, 2, 3, 6, 7], nb=[0, 2],
            send=[[0, 1], [1, 3]], recv=[[4, 5], [6, 7]]),
    2: dict(n=4, nnz=20, rp=[0, 5, 9, 16, 20], ci=[4, 5, 0, 2, 3, 7, 1, 2, 3, 5, 6, 7, 0, 1, 2, 3, 0, 1, 2, 3], nb=[0, 1],
            send=[[0, 2], [1, 2]], recv=[[4, 5], [6, 7]]),
}


@pytest.fixture(scope="module")
def mtx_file(tmp_path_factory):
    p = tmp_path_factory.mktemp("sys") / "example12.mtx"
    lines = ["%%MatrixMarket matrix coordinate real general", "%%AMGX rhs", f"12 12 {len(CI)}"]
    for i in range(12):
        for k in range(RP[i], RP[i + 1]):
            lines.append(f"{i + 1} {CI[k] + 1} {100 * (i + 1) + CI[k] + 1}")      # value encodes (row, col)
    lines += [str(float(i + 1)) for i in range(12)]
    p.write_text("\n".join(lines) + "\n")
    return str(p)


def call(lib, rank, world, filename, pv=None, sizes=None, nparts=0, want_local=True, want_global=True):
    n, nnz, bx, by, nn = (C.c_int() for _ in range(5))
    rp, cl, nb, ssz, rsz = (C.POINTER(C.c_int)() for _ in range(5))
    cg = C.POINTER(C.c_int64)()
    data, diag, rhs, sol = (C.c_void_p() for _ in range(4))
    sm, rm = C.POINTER(C.POINTER(C.c_int))(), C.POINTER(C.POINTER(C.c_int))()
    pva = (C.c_int * len(pv))(*pv) if pv is not None else None
    sza = (C.c_int * len(sizes))(*sizes) if sizes is not None else None
    f = lib.AMGXB200_read_system_partition
    f.restype = C.c_int
    f.argtypes = None
    rc = f(rank, world, 8193, filename.encode(), nparts, sza, len(pv) if pv is not None else 0, pva, C.byref(n), C.byref(nnz), C.byref(bx), C.byref(by),
           C.byref(rp), C.byref(cl) if want_local else None, C.byref(cg) if want_global else None, C.byref(data), C.byref(diag), C.byref(rhs),
           C.byref(sol), C.byref(nn), C.byref(nb), C.byref(ssz), C.byref(sm), C.byref(rsz), C.byref(rm))
    if rc != 0:
        return rc, None
    out = dict(n=n.value, nnz=nnz.value, bx=bx.value, by=by.value, rp=[rp[i] for i in range(n.value + 1)],
               data=np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_double)), (max(nnz.value, 1),))[: nnz.value].copy(),
               rhs=np.ctypeslib.as_array(C.cast(rhs, C.POINTER(C.c_double)), (max(n.value, 1),))[: n.value].copy(),
               sol=np.ctypeslib.as_array(C.cast(sol, C.POINTER(C.c_double)), (max(n.value, 1),))[: n.value].copy())
    if want_global:
        out["cg"] = [cg[i] for i in range(nnz.value)]
    if want_local:
        out["ci"] = [cl[i] for i in range(nnz.value)]
        out["nb"] = [nb[q] for q in range(nn.value)]
        out["send"] = [[sm[q][k] for k in range(ssz[q])] for q in range(nn.value)]
        out["recv"] = [[rm[q][k] for k in range(rsz[q])] for q in range(nn.value)]
        lib.AMGX_free_system_maps_one_ring.argtypes = None
        lib.AMGX_free_system_maps_one_ring(rp, cl, data, diag, rhs, sol, nn.value, nb, ssz, sm, rsz, rm)
    return 0, out


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_maps_one_ring_matches_the_reference_worked_example(rank, mtx_file):
    lib = capi.load_library()
    rc, o = call(lib, rank, 3, mtx_file, pv=PV)
    assert rc == 0
    e = EXPECT[rank]
    assert (o["n"], o["nnz"], o["bx"], o["by"]) == (e["n"], e["nnz"], 1, 1)
    assert o["rp"] == e["rp"] and o["ci"] == e["ci"]
    assert o["nb"] == e["nb"] and o["send"] == e["send"] and o["recv"] == e["recv"]
    # read_system_global: the same rows with their ORIGINAL global column ids; values / rhs follow the rows
    lo = 4 * rank
    assert o["cg"] == CI[RP[lo]:RP[lo + 4]]
    rows = np.repeat(np.arange(lo, lo + 4), np.diff(RP[lo:lo + 5]))
    assert np.array_equal(o["data"], 100.0 * (rows + 1) + np.array(o["cg"]) + 1)
    assert np.array_equal(o["rhs"], np.arange(lo + 1, lo + 5, dtype=float)) and not o["sol"].any()


def test_default_and_sized_partitions_and_scattered_vector(mtx_file):
    lib = capi.load_library()
    # no partition info: equal contiguous blocks p*n/ranks
    for r, (lo, hi) in enumerate([(0, 6), (6, 12)]):
        rc, o = call(lib, r, 2, mtx_file)
        assert rc == 0 and o["n"] == hi - lo and o["cg"] == CI[RP[lo]:RP[hi]]
    # partition_sizes: contiguous blocks of the given sizes; 4 partitions on 2 ranks = two consecutive partitions per rank
    rc, o = call(lib, 1, 2, mtx_file, sizes=[2, 3, 3, 4], nparts=4)
    assert rc == 0 and o["n"] == 7 and o["cg"] == CI[RP[5]:RP[12]]
    # scattered vector: rows kept in increasing global id, local numbering consistent with the maps
    pv = [1, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 1]
    views = [call(lib, r, 2, mtx_file, pv=pv)[1] for r in range(2)]
    mine = [[g for g in range(12) if pv[g] == r] for r in range(2)]
    for r in range(2):
        o, other = views[r], views[1 - r]
        assert o["n"] == len(mine[r]) and o["nb"] == [1 - r]
        assert o["cg"] == [c for g in mine[r] for c in CI[RP[g]:RP[g + 1]]]
        # a halo column's global id = the row the neighbour sends for it
        halo_gid = {h: mine[1 - r][other["send"][0][k]] for k, h in enumerate(o["recv"][0])}
        for lc, gc in zip(o["ci"], o["cg"]):
            assert (mine[r][lc] if lc < o["n"] else halo_gid[lc]) == gc
    # errors: wrong vector length, sizes that do not add up
    assert call(lib, 0, 2, mtx_file, pv=[0, 1, 0])[0] != 0
    assert call(lib, 0, 2, mtx_file, sizes=[5, 5], nparts=2)[0] != 0


def test_binary_system_file_is_detected_and_read(tmp_path):
    """the reference's "%%NVAMGBinary" layout (src/matrix_io.cu:267-405): header, 9 uint32 flags, int32 offsets / columns, float64 values
    (+ diagonal blocks), rhs, solution"""
    import struct
    lib = capi.load_library()
    rp = np.array(RP, np.int32)
    ci = np.array(CI, np.int32)
    va = np.arange(1, len(CI) + 1, dtype=np.float64) * 0.5
    rhs = np.linspace(1, 2, 12)
    sol = np.linspace(-1, 1, 12)
    p = tmp_path / "sys.bin"
    with open(p, "wb") as f:
        f.write(b"%%NVAMGBinary\n")
        f.write(struct.pack("<9I", 1, 1, 1, 0, 0, 1, 1, 12, len(CI)))
        f.write(rp.tobytes() + ci.tobytes() + va.tobytes() + rhs.tobytes() + sol.tobytes())
    rc, o = call(lib, 0, 1, str(p))
    assert rc == 0 and o["n"] == 12 and o["nnz"] == len(CI) and o["rp"] == RP and o["cg"] == CI and o["ci"] == CI
    assert np.array_equal(o["data"], va) and np.array_equal(o["rhs"], rhs) and np.array_equal(o["sol"], sol)
    # partitioned read of the same binary file
    rc, o1 = call(lib, 1, 3, str(p), pv=PV)
    assert rc == 0 and o1["ci"] == EXPECT[1]["ci"] and np.array_equal(o1["rhs"], rhs[4:8])
    # a truncated file fails loudly
    q = tmp_path / "short.bin"
    q.write_bytes(p.read_bytes()[:200])
    assert call(lib, 0, 1, str(q))[0] != 0


def _write(path, header, entries, tail_lines):
    path.write_text("\n".join(["%%MatrixMarket matrix coordinate real general", header, f"12 12 {len(entries)}"] + entries + tail_lines) + "\n")
    return str(path)


@pytest.mark.parametrize("with_lengths", [True, False])
def test_reference_file_format_sections(tmp_path, with_lengths):
    """What the reference's own writer emits after the entries (src/matrix_io.cu:222-258) and its reader expects (src/readers.cu:1290-1406):
    with "diagonal" one line per row holding the diagonal block (which is then not among the entries), and a LENGTH line in front of
    the rhs and of the solution.  Files without the length lines (older files of this engine, hand-written ones) still read the same."""
    lib = capi.load_library()
    off = [(i, CI[k]) for i in range(12) for k in range(RP[i], RP[i + 1]) if CI[k] != i]
    entries = [f"{i + 1} {j + 1} {100 * (i + 1) + j + 1}" for i, j in off]
    diag = [f"{1000.0 + i!r} " for i in range(12)]
    rhs, sol = [str(float(i + 1)) for i in range(12)], [str(0.5 * i) for i in range(12)]
    tail = diag + (["12"] if with_lengths else []) + rhs + (["12"] if with_lengths else []) + sol
    fn = _write(tmp_path / "ref_format.mtx", "%%NVAMG 1 1 diagonal rhs solution", entries, tail)
    rc, got = call(lib, 0, 1, fn, want_local=False)
    assert rc == 0
    assert got["n"] == 12 and np.array_equal(got["rhs"], np.arange(1.0, 13.0)) and np.array_equal(got["sol"], 0.5 * np.arange(12))
    # the external diagonal is handed back separately or merged, depending on the hook; either way every value must be found once
    vals = sorted(got["data"].tolist())
    assert vals[: len(off)] == sorted(100.0 * (i + 1) + j + 1 for i, j in off)
    # the same system with the diagonal INLINE (no "diagonal" keyword) and no solution section
    inline = [f"{i + 1} {CI[k] + 1} {100 * (i + 1) + CI[k] + 1}" for i in range(12) for k in range(RP[i], RP[i + 1])]
    fn2 = _write(tmp_path / "inline.mtx", "%%AMGX rhs", inline, (["12"] if with_lengths else []) + rhs)
    rc, got2 = call(lib, 0, 1, fn2, want_local=False)
    assert rc == 0 and got2["nnz"] == len(CI) and np.array_equal(got2["rhs"], np.arange(1.0, 13.0))
    # a wrong number of trailing values is an error, not a silent misread
    fn3 = _write(tmp_path / "short.mtx", "%%AMGX rhs", inline, rhs[:-2])
    rc, _ = call(lib, 0, 1, fn3, want_local=False)
    assert rc != 0


@pytest.mark.parametrize("writer", ["matrixmarket", "binary"])
@pytest.mark.parametrize("block,ext_diag", [(1, 0), (1, 1), (2, 0), (2, 1)])
def test_writer_reader_round_trip_on_the_host(tmp_path, writer, block, ext_diag):
    """AMGX_write_system's file writer and AMGX_read_system's reader, both through their resource-free hooks: matrix (scalar and 2x2
    blocks, diagonal inside or outside the CSR structure), right-hand side and solution survive a round trip bit for bit in both formats"""
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    n = 12
    rows = [(i, CI[k]) for i in range(n) for k in range(RP[i], RP[i + 1]) if not (ext_diag and CI[k] == i)]
    rp = np.zeros(n + 1, np.int32)
    for i, _ in rows:
        rp[i + 1] += 1
    rp = np.cumsum(rp).astype(np.int32)
    ci = np.array([j for _, j in rows], np.int32)
    bsq = block * block
    va = rng.standard_normal((len(rows) + (n if ext_diag else 0)) * bsq)
    b, x = rng.standard_normal(n * block), rng.standard_normal(n * block)
    fn = str(tmp_path / f"rt_{writer}_{block}_{ext_diag}.dat").encode()
    f = lib.AMGXB200_write_system_host
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert f(fn, writer.encode(), n, len(rows), block, block, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ext_diag, b.ctypes.data, x.ctypes.data) == 0
    # read back: one partition, global columns
    nn, nnz, bx, by, nnb = (C.c_int() for _ in range(5))
    prp = C.POINTER(C.c_int)()
    cg = C.POINTER(C.c_int64)()
    data, diag, rhs, sol = (C.c_void_p() for _ in range(4))
    g = lib.AMGXB200_read_system_partition
    g.restype = C.c_int
    g.argtypes = None
    rc = g(0, 1, 8193, fn, 0, None, 0, None, C.byref(nn), C.byref(nnz), C.byref(bx), C.byref(by), C.byref(prp), None, C.byref(cg), C.byref(data), C.byref(diag),
           C.byref(rhs), C.byref(sol), C.byref(nnb), None, None, None, None, None)
    assert rc == 0
    assert (nn.value, nnz.value, bx.value, by.value) == (n, len(rows), block, block)
    assert [prp[i] for i in range(n + 1)] == rp.tolist() and [cg[k] for k in range(nnz.value)] == ci.tolist()
    arr = lambda p, m: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), (m,)).copy()
    assert np.array_equal(arr(data, len(rows) * bsq), va[: len(rows) * bsq])
    if ext_diag:
        assert diag.value and np.array_equal(arr(diag, n * bsq), va[len(rows) * bsq:])
    assert np.array_equal(arr(rhs, n * block), b) and np.array_equal(arr(sol, n * block), x)


def test_one_row_system_with_diagonal_section_is_not_misread(tmp_path):
    """with n * block size == number of length lines the value COUNT cannot tell the layouts apart (1-row system, 'diagonal' +
    'rhs': diagonal section + rhs without a length line has as many values as inline diagonal + length line + rhs).  The reader checks that
    a claimed length line really holds the vector length."""
    lib = capi.load_library()
    base = ["%%MatrixMarket matrix coordinate real general"]
    # reference layout: no entries, one diagonal line, a length line, the rhs
    p1 = tmp_path / "one_ref.mtx"
    p1.write_text("\n".join(base + ["%%NVAMG 1 1 diagonal rhs", "1 1 0", "7.5", "1", "3.25"]) + "\n")
    rc, got = call(lib, 0, 1, str(p1), want_local=False)
    assert rc == 0 and got["n"] == 1 and got["rhs"][0] == 3.25
    # legacy layout: diagonal section + rhs WITHOUT the length line; the rhs value 1.0 must not be taken for a length line
    p2 = tmp_path / "one_legacy.mtx"
    p2.write_text("\n".join(base + ["%%NVAMG 1 1 diagonal rhs", "1 1 0", "7.5", "4.0"]) + "\n")
    rc, got = call(lib, 0, 1, str(p2), want_local=False)
    assert rc == 0 and got["rhs"][0] == 4.0
