"""Worker for the CPU (gloo) multi-process test of the host-side partition logic: each rank owns a z-slab of a
7-pt Poisson problem, plans its partition through the C-ABI planner (pure host code), exchanges halos over
gloo exactly as the engine does over NCCL (pack by send map -> send/recv -> halo tail) and checks that the
distributed SpMV equals the global one bit for bit."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi, gallery  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    orc.set_num_threads(1)
    nx, ny, nzl = 7, 5, 4
    nz = nzl * world
    rp, ci, va = gallery.poisson7pt(nx, ny, nz)
    n_global = rp.shape[0] - 1
    lib = capi.load_library()
    if os.environ.get("AMGXB_TEST_PARTITION") == "vector":
        return partition_vector_case(lib, rank, world, rp, ci, va)
    if os.environ.get("AMGXB_TEST_PARTITION") == "maps":
        return comm_maps_case(lib, rank, world, rp, ci, va, nx * ny * nzl)
    offsets = np.array([nx * ny * nzl * r for r in range(world + 1)], np.int64)
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    lrp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
    lci = ci[rp[lo]:rp[hi]].astype(np.int64)
    lva = va[rp[lo]:rp[hi]]
    plan = capi.PartitionPlan()
    rc = lib.AMGXB200_partition_plan_create(C.byref(plan), rank, world, offsets.ctypes.data, hi - lo, lci.shape[0], lrp.ctypes.data, lci.ctypes.data)
    assert rc == 0, rc
    n, nh, nn, n_int = plan.n_owned, plan.n_halo, plan.num_neighbors, plan.n_interior
    neighbors = np.ctypeslib.as_array(plan.neighbors, (nn,)).copy()
    send_off = np.ctypeslib.as_array(plan.send_offsets, (nn + 1,)).copy()
    send_maps = np.ctypeslib.as_array(plan.send_maps, (max(send_off[-1], 1),)).copy()[: send_off[-1]]
    halo_off = np.ctypeslib.as_array(plan.halo_offsets, (nn + 1,)).copy()
    perm = np.ctypeslib.as_array(plan.perm_old_to_new, (n,)).copy()
    lcols = np.ctypeslib.as_array(plan.local_cols, (lci.shape[0],)).copy()
    halo_global = np.ctypeslib.as_array(plan.halo_global, (max(nh, 1),)).copy()[:nh]
    lib.AMGXB200_partition_plan_free(C.byref(plan))
    # structure checks
    expected_nb = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
    assert list(neighbors) == expected_nb, (neighbors, expected_nb)
    assert nh == nx * ny * len(expected_nb) and n_int == n - nx * ny * len(expected_nb)
    assert sorted(perm) == list(range(n)) and np.all(np.sort(perm[: 0]) == [])
    # local matrix in the renumbered row order
    inv = np.argsort(perm)
    rows_len = np.diff(lrp)[inv]
    nrp = np.zeros(n + 1, np.int32)
    np.cumsum(rows_len, out=nrp[1:])
    nci = np.concatenate([lcols[lrp[i]:lrp[i + 1]] for i in inv]).astype(np.int32)
    nva = np.concatenate([lva[lrp[i]:lrp[i + 1]] for i in inv])
    assert nci[: nrp[n_int]].max(initial=-1) < n           # interior rows never touch the halo
    # distributed x: owned part permuted, halo tail received from the neighbours
    xg = np.random.default_rng(42).standard_normal(n_global)
    x = np.zeros(n + nh)
    x[perm] = xg[lo:hi]
    reqs, bufs = [], []
    for q, nb in enumerate(neighbors):
        sb = torch.from_numpy(x[send_maps[send_off[q]:send_off[q + 1]]].copy())
        rb = torch.empty(int(halo_off[q + 1] - halo_off[q]), dtype=torch.float64)
        reqs.append(dist.isend(sb, int(nb)))
        reqs.append(dist.irecv(rb, int(nb)))
        bufs.append((q, sb, rb))
    for r in reqs:
        r.wait()
    for q, _, rb in bufs:
        x[n + halo_off[q]: n + halo_off[q + 1]] = rb.numpy()
    assert np.array_equal(x[n:], xg[halo_global])                  # halo values are the right global entries
    y = np.empty(n)
    orc.lib().orc_spmv(n, nrp.ctypes.data_as(C.c_void_p), nci.ctypes.data_as(C.c_void_p), nva.ctypes.data_as(C.c_void_p),
                       x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    yg = orc.spmv(rp, ci, va, xg)
    assert np.array_equal(y[perm], yg[lo:hi]), "distributed SpMV differs from the global one"
    # global dot through all-reduce == sum of local dots (order differs: tolerance)
    t = torch.tensor([float(np.dot(y, y))], dtype=torch.float64)
    dist.all_reduce(t)
    assert abs(t.item() - float(np.dot(yg, yg))) <= 1e-12 * float(np.dot(yg, yg))
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK")
    dist.destroy_process_group()


def comm_maps_case(lib, rank, world, rp, ci, va, nloc):
    """AMGX_matrix_comm_from_maps_one_ring semantics (examples/amgx_mpi_capi_agg.c:386-420): every rank holds its rows in LOCAL
    numbering plus neighbours / send_maps / recv_maps.  The engine exchanges the senders' global row ids and calls the host helper
    tested here; the result must be the true global columns, and the usual planner must then reproduce the global SpMV."""
    n_global = rp.shape[0] - 1
    offsets = np.array([nloc * r for r in range(world + 1)], np.int64)
    owner = lambda g: int(np.searchsorted(offsets, g, side="right") - 1)

    def local_view(r):
        """what AMGX_read_system_maps_one_ring would hand rank r; halo columns numbered in a deliberately scrambled order"""
        lo, hi = int(offsets[r]), int(offsets[r + 1])
        cols = ci[rp[lo]:rp[hi]]
        halo = np.unique(cols[(cols < lo) | (cols >= hi)])
        halo = np.random.default_rng(100 + r).permutation(halo)              # arbitrary halo numbering
        hid = {int(g): hi - lo + k for k, g in enumerate(halo)}
        lcols = np.array([c - lo if lo <= c < hi else hid[int(c)] for c in cols], np.int32)
        nbrs = sorted({owner(g) for g in halo})
        recv = {q: np.array([hid[int(g)] for g in halo if owner(g) == q], np.int32) for q in nbrs}
        recv_g = {q: np.array([int(g) for g in halo if owner(g) == q], np.int64) for q in nbrs}
        return lo, hi, lcols, nbrs, recv, recv_g

    lo, hi, lcols, nbrs, recv, recv_g = local_view(rank)
    n = hi - lo
    # send_maps[q][k] = my local row whose value lands in q's recv_maps[me][k] (the pairing rule of the API)
    send = {q: (local_view(q)[5][rank] - lo).astype(np.int32) for q in nbrs}
    # ---- what the engine does over NCCL, here over gloo: send the global ids of my send rows, receive my halo's global ids
    reqs, got = [], {}
    for q in nbrs:
        sb = torch.from_numpy((send[q].astype(np.int64) + lo).copy())
        rb = torch.empty(recv[q].shape[0], dtype=torch.int64)
        reqs += [dist.isend(sb, q), dist.irecv(rb, q)]
        got[q] = (sb, rb)
    for r in reqs:
        r.wait()
    nn = len(nbrs)
    rsz = (C.c_int * max(nn, 1))(*[recv[q].shape[0] for q in nbrs])
    rmaps = (C.c_void_p * max(nn, 1))(*[recv[q].ctypes.data for q in nbrs])
    rglob_np = [got[q][1].numpy().copy() for q in nbrs]
    rglob = (C.c_void_p * max(nn, 1))(*[a.ctypes.data for a in rglob_np])
    out = np.empty(lcols.shape[0], np.int64)
    rc = lib.AMGXB200_comm_maps_to_global_cols(n, lcols.shape[0], lcols.ctypes.data, lo, nn, rsz, rmaps, rglob, out.ctypes.data)
    assert rc == 0, rc
    assert np.array_equal(out, ci[rp[lo]:rp[hi]].astype(np.int64)), "comm maps did not reproduce the global columns"
    # a map that misses a referenced halo column must be rejected
    if nn:
        short = (C.c_int * nn)(*([recv[nbrs[0]].shape[0] - 1] + [recv[q].shape[0] for q in nbrs[1:]]))
        scratch = out.copy()
        assert lib.AMGXB200_comm_maps_to_global_cols(n, lcols.shape[0], lcols.ctypes.data, lo, nn, short, rmaps, rglob, scratch.ctypes.data) != 0
    # ---- the converted matrix goes through the usual planner: distributed SpMV == global SpMV
    lrp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
    plan = capi.PartitionPlan()
    assert lib.AMGXB200_partition_plan_create(C.byref(plan), rank, world, offsets.ctypes.data, n, out.shape[0], lrp.ctypes.data, out.ctypes.data) == 0
    assert plan.n_halo == sum(recv[q].shape[0] for q in nbrs) and plan.num_neighbors == nn
    lib.AMGXB200_partition_plan_free(C.byref(plan))
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK")
    dist.destroy_process_group()


def partition_vector_case(lib, rank, world, rp, ci, va):
    """an arbitrary (scattered) partition vector: rows are renumbered into the contiguous layout of the reference's ipartition_map
    (AMGXB200_partition_vector_to_contiguous), then planned / exchanged / multiplied exactly like a contiguous partition"""
    n_global = rp.shape[0] - 1
    pv = np.random.default_rng(7).integers(0, world, n_global).astype(np.int32)      # same on every rank
    offsets = np.zeros(world + 1, np.int64)
    new_global = np.zeros(n_global, np.int64)
    rc = lib.AMGXB200_partition_vector_to_contiguous(n_global, world, pv.ctypes.data, offsets.ctypes.data, new_global.ctypes.data)
    assert rc == 0, rc
    # the map is a permutation that sends rank r's rows, in increasing global id, onto [offsets[r], offsets[r+1])
    assert sorted(new_global) == list(range(n_global))
    for r in range(world):
        mine_r = np.nonzero(pv == r)[0]
        assert np.array_equal(new_global[mine_r], np.arange(offsets[r], offsets[r + 1]))
    bad = pv.copy()
    bad[3] = world
    off2 = offsets.copy()
    assert lib.AMGXB200_partition_vector_to_contiguous(n_global, world, bad.ctypes.data, off2.ctypes.data, None) != 0
    mine = np.nonzero(pv == rank)[0]                     # my rows, increasing global id: the order callers upload them in
    n = mine.shape[0]
    lens = (rp[mine + 1] - rp[mine]).astype(np.int32)
    lrp = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=lrp[1:])
    lci = np.concatenate([new_global[ci[rp[g]:rp[g + 1]]] for g in mine]).astype(np.int64)
    lva = np.concatenate([va[rp[g]:rp[g + 1]] for g in mine])
    plan = capi.PartitionPlan()
    rc = lib.AMGXB200_partition_plan_create(C.byref(plan), rank, world, offsets.ctypes.data, n, lci.shape[0], lrp.ctypes.data, lci.ctypes.data)
    assert rc == 0, rc
    nh, nn = plan.n_halo, plan.num_neighbors
    neighbors = np.ctypeslib.as_array(plan.neighbors, (max(nn, 1),)).copy()[:nn]
    send_off = np.ctypeslib.as_array(plan.send_offsets, (nn + 1,)).copy()
    send_maps = np.ctypeslib.as_array(plan.send_maps, (max(send_off[-1], 1),)).copy()[: send_off[-1]]
    halo_off = np.ctypeslib.as_array(plan.halo_offsets, (nn + 1,)).copy()
    perm = np.ctypeslib.as_array(plan.perm_old_to_new, (max(n, 1),)).copy()[:n]
    lcols = np.ctypeslib.as_array(plan.local_cols, (max(lci.shape[0], 1),)).copy()[: lci.shape[0]]
    halo_global = np.ctypeslib.as_array(plan.halo_global, (max(nh, 1),)).copy()[:nh]
    lib.AMGXB200_partition_plan_free(C.byref(plan))
    inv = np.argsort(perm)
    nrp = np.zeros(n + 1, np.int32)
    np.cumsum(np.diff(lrp)[inv], out=nrp[1:])
    nci = np.concatenate([lcols[lrp[i]:lrp[i + 1]] for i in inv]).astype(np.int32)
    nva = np.concatenate([lva[lrp[i]:lrp[i + 1]] for i in inv])
    xg = np.random.default_rng(42).standard_normal(n_global)        # indexed by ORIGINAL global id
    x = np.zeros(n + nh)
    x[perm] = xg[mine]
    reqs, bufs = [], []
    for q, nb in enumerate(neighbors):
        sb = torch.from_numpy(x[send_maps[send_off[q]:send_off[q + 1]]].copy())
        rb = torch.empty(int(halo_off[q + 1] - halo_off[q]), dtype=torch.float64)
        reqs.append(dist.isend(sb, int(nb)))
        reqs.append(dist.irecv(rb, int(nb)))
        bufs.append((q, rb))
    for r in reqs:
        r.wait()
    for q, rb in bufs:
        x[n + halo_off[q]: n + halo_off[q + 1]] = rb.numpy()
    old_of_new = np.argsort(new_global)                              # contiguous id -> original global id
    assert np.array_equal(x[n:], xg[old_of_new[halo_global]])
    y = np.empty(n)
    orc.lib().orc_spmv(n, nrp.ctypes.data_as(C.c_void_p), nci.ctypes.data_as(C.c_void_p), nva.ctypes.data_as(C.c_void_p),
                       x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    yg = orc.spmv(rp, ci, va, xg)
    assert np.array_equal(y[perm], yg[mine]), "distributed SpMV (partition vector) differs from the global one"
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
