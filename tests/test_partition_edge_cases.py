"""CPU: edge cases of the host-side partition planner (csrc/partition.cpp) called directly through the C-ABI, no process group: an empty
rank, rows without entries, a rank without neighbours, a partition vector that leaves a rank empty."""
import ctypes as C

import numpy as np
import pytest

from amgx_b200 import capi, gallery


def _plan(lib, rank, world, offsets, rp, ci):
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    lrp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
    lci = np.ascontiguousarray(ci[rp[lo]:rp[hi]].astype(np.int64))
    if lci.shape[0] == 0:
        lci = np.zeros(1, np.int64)
    plan = capi.PartitionPlan()
    rc = lib.AMGXB200_partition_plan_create(C.byref(plan), rank, world, offsets.ctypes.data, hi - lo, int(lrp[-1]), lrp.ctypes.data, lci.ctypes.data)
    assert rc == 0, rc
    n, nh, nn = plan.n_owned, plan.n_halo, plan.num_neighbors
    out = dict(n=n, nh=nh, nn=nn, n_int=plan.n_interior,
               neighbors=np.ctypeslib.as_array(plan.neighbors, (max(nn, 1),)).copy()[:nn],
               send_off=np.ctypeslib.as_array(plan.send_offsets, (nn + 1,)).copy(),
               halo_off=np.ctypeslib.as_array(plan.halo_offsets, (nn + 1,)).copy(),
               perm=np.ctypeslib.as_array(plan.perm_old_to_new, (max(n, 1),)).copy()[:n],
               lcols=np.ctypeslib.as_array(plan.local_cols, (max(int(lrp[-1]), 1),)).copy()[:int(lrp[-1])],
               halo_global=np.ctypeslib.as_array(plan.halo_global, (max(nh, 1),)).copy()[:nh], lrp=lrp, lo=lo, hi=hi)
    out["send_maps"] = np.ctypeslib.as_array(plan.send_maps, (max(int(out["send_off"][-1]), 1),)).copy()[:int(out["send_off"][-1])]
    lib.AMGXB200_partition_plan_free(C.byref(plan))
    return out


def _check(p, ci, rp, offsets):
    n = p["n"]
    assert sorted(p["perm"]) == list(range(n))
    assert p["lcols"].shape[0] == int(p["lrp"][-1])
    if p["lcols"].shape[0]:
        assert p["lcols"].min() >= 0 and p["lcols"].max() < n + p["nh"]
    # every local column maps back to the global column it came from
    gl = ci[rp[p["lo"]]:rp[p["hi"]]]
    inv = np.empty(max(n, 1), np.int64)
    inv[p["perm"]] = np.arange(n)
    back = np.where(p["lcols"] < n, p["lo"] + inv[np.minimum(p["lcols"], max(n - 1, 0))], 0)
    halo = p["lcols"] >= n
    back[halo] = p["halo_global"][p["lcols"][halo] - n]
    assert np.array_equal(back, gl)
    # halo ids are grouped by owner in neighbour order
    owner = np.searchsorted(offsets, p["halo_global"], side="right") - 1
    for q, nb in enumerate(p["neighbors"]):
        assert np.all(owner[p["halo_off"][q]:p["halo_off"][q + 1]] == nb)
    assert p["halo_off"][-1] == p["nh"]
    # interior rows (new index < n_int) reference no halo column
    for old in range(n):
        cols = p["lcols"][p["lrp"][old]:p["lrp"][old + 1]]
        if p["perm"][old] < p["n_int"]:
            assert not np.any(cols >= n)


def test_empty_rank_and_rank_without_neighbours():
    lib = capi.load_library()
    rp, ci, va = gallery.poisson7pt(5, 4, 6)
    ng = rp.shape[0] - 1
    # rank 1 owns nothing; ranks 0 and 2 share the cut
    offsets = np.array([0, 60, 60, ng], np.int64)
    plans = [_plan(lib, r, 3, offsets, rp, ci) for r in range(3)]
    assert plans[1]["n"] == 0 and plans[1]["nh"] == 0 and plans[1]["nn"] == 0
    assert list(plans[0]["neighbors"]) == [2] and list(plans[2]["neighbors"]) == [0]
    for p in plans:
        _check(p, ci, rp, offsets)
    # what rank 0 sends to rank 2 is what rank 2 expects from rank 0, in the same order (global ids)
    p0, p2 = plans[0], plans[2]
    inv0 = np.empty(p0["n"], np.int64)
    inv0[p0["perm"]] = np.arange(p0["n"])
    sent = p0["lo"] + inv0[p0["send_maps"][p0["send_off"][0]:p0["send_off"][1]]]
    assert np.array_equal(sent, p2["halo_global"][p2["halo_off"][0]:p2["halo_off"][1]])
    # block-diagonal matrix: nobody has neighbours
    import scipy.sparse as sp
    B = sp.block_diag([gallery.to_scipy(*gallery.poisson7pt(3))] * 2).tocsr()
    brp, bci = B.indptr.astype(np.int32), B.indices.astype(np.int32)
    off2 = np.array([0, 27, 54], np.int64)
    for r in range(2):
        p = _plan(lib, r, 2, off2, brp, bci)
        assert p["nn"] == 0 and p["nh"] == 0 and p["n_int"] == 27
        _check(p, bci, brp, off2)


def test_rows_without_entries():
    lib = capi.load_library()
    rp, ci, va = gallery.poisson7pt(4, 3, 4)
    A = gallery.to_scipy(rp, ci, va).tolil()
    for i in (0, 7, 23, 24, 47):          # empty rows on both sides of the cut
        A.rows[i], A.data[i] = [], []
    A = A.tocsr()
    rp2, ci2 = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    offsets = np.array([0, 24, 48], np.int64)
    for r in range(2):
        p = _plan(lib, r, 2, offsets, rp2, ci2)
        _check(p, ci2, rp2, offsets)


def test_partition_vector_with_an_empty_rank():
    lib = capi.load_library()
    pv = np.array([2, 0, 0, 2, 0, 2, 2, 0], np.int32)        # rank 1 owns nothing
    offsets = np.zeros(4, np.int64)
    newg = np.zeros(8, np.int64)
    rc = lib.AMGXB200_partition_vector_to_contiguous(8, 3, pv.ctypes.data, offsets.ctypes.data, newg.ctypes.data)
    assert rc == 0
    assert list(offsets) == [0, 4, 4, 8]
    assert list(newg) == [4, 0, 1, 5, 2, 6, 7, 3]
    bad = np.array([0, 3], np.int32)
    assert lib.AMGXB200_partition_vector_to_contiguous(2, 3, bad.ctypes.data, offsets.ctypes.data, newg.ctypes.data) != 0
