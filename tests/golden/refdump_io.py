"""IO helpers for the reference dump harness (oracle/ref_build/ref_dump.cu): write system.bin, read out.bin."""
from __future__ import annotations

import struct

import numpy as np


def write_system(path, rp, ci, va, rhs, diag=None, x0=None, block=(1, 1)):
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    va = np.ascontiguousarray(va, np.float64)
    n, nnz = rp.shape[0] - 1, ci.shape[0]
    with open(path, "wb") as f:
        f.write(np.array([n, nnz, block[0], block[1], int(diag is not None), int(x0 is not None)], np.int32).tobytes())
        f.write(rp.tobytes())
        f.write(ci.tobytes())
        f.write(va.tobytes())
        if diag is not None:
            f.write(np.ascontiguousarray(diag, np.float64).tobytes())
        f.write(np.ascontiguousarray(rhs, np.float64).tobytes())
        if x0 is not None:
            f.write(np.ascontiguousarray(x0, np.float64).tobytes())


def read_dump(path) -> dict:
    out = {}
    dts = {b"i": np.int32, b"d": np.float64, b"f": np.float32, b"c": np.uint8}
    with open(path, "rb") as f:
        data = f.read()
    p = 0
    while p < len(data):
        (nl,) = struct.unpack_from("<I", data, p)
        p += 4
        name = data[p:p + nl].decode()
        p += nl
        dt = data[p:p + 1]
        p += 1
        (cnt,) = struct.unpack_from("<Q", data, p)
        p += 8
        dtype = np.dtype(dts[dt])
        arr = np.frombuffer(data, dtype=dtype, count=cnt, offset=p).copy()
        p += cnt * dtype.itemsize
        if dt == b"c":
            out[name] = arr.tobytes().decode(errors="replace")
        else:
            out[name] = arr
    return out
