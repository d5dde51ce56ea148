"""Config 5 (4x4 blocks, AMG + MULTICOLOR_DILU, mixed precision) weak scaling over z-slabs: one process per GPU (torchrun)."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi, gallery  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
mode = sys.argv[1] if len(sys.argv) > 1 else "dDFI"
nx = ny = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nzl = int(sys.argv[3]) if len(sys.argv) > 3 else 80
outer = sys.argv[4] if len(sys.argv) > 4 else "PCG"
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
capi.initialize()
capi.register_print_callback(None)
amg = {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
       "matrix_coloring_scheme": "MIN_MAX", "max_uncolored_percentage": 0.15, "smoother": "MULTICOLOR_DILU", "relaxation_factor": 0.9,
       "presweeps": 1, "postsweeps": 1, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER"}
tol = 1e-6
if outer == "AMG":
    cfgd = {"config_version": 2, "solver": dict(amg, scope="main", max_iters=100, monitor_residual=1, store_res_history=1, convergence="RELATIVE_INI",
                                                tolerance=tol, norm="L2")}
else:
    cfgd = {"config_version": 2, "solver": {"scope": "main", "solver": outer, "max_iters": 100, "gmres_n_restart": 20, "monitor_residual": 1,
                                            "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": tol, "norm": "L2",
                                            "preconditioner": dict(amg, max_iters=1, monitor_residual=0)}}
cfg = capi.Config(cfgd)
comm = None
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = capi.AMGXB200_comm(rank, world, idt.cpu().numpy().tobytes())
rsc = capi.Resources(cfg, device=lr, comm=comm) if comm is not None else capi.Resources(cfg, device=lr)
nz = nzl * world
t = time.time()
lrp, lci, lva = gallery.block_elasticity_slab(nx, ny, nz, nzl * rank, nzl * (rank + 1), dtype=np.float32 if mode == "dDFI" else np.float64)
tgen = time.time() - t
n = lrp.shape[0] - 1
ng = nx * ny * nz
lib = capi.load_library()
A = capi.Matrix(rsc, mode)
if world > 1:
    offsets = np.array([nx * ny * nzl * r for r in range(world + 1)], np.int64)
    dh = C.c_void_p()
    assert lib.AMGX_distribution_create(C.byref(dh), cfg.h) == 0
    assert lib.AMGX_distribution_set_partition_data(dh, 1, offsets.ctypes.data) == 0
    rc = lib.AMGX_matrix_upload_distributed(A.h, ng, n, lci.shape[0], 4, 4, lrp.ctypes.data, lci.ctypes.data, lva.ctypes.data, None, dh)
    assert rc == 0, rc
    lib.AMGX_distribution_destroy(dh)
else:
    A.upload(lrp, lci.astype(np.int32), lva, block_dims=(4, 4))
b, x = capi.Vector(rsc, mode), capi.Vector(rsc, mode)
if world > 1:
    b.bind(A)
    x.bind(A)
b.upload(np.ones(n * 4), block_dim=4)
x.set_zero(n, 4)
slv = capi.Solver(rsc, cfg, mode)
t = time.time()
slv.setup(A)
tsetup = time.time() - t
best = 1e30
for _ in range(3):
    x.set_zero(n, 4)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    slv.solve(b, x, zero_initial_guess=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    best = min(best, time.time() - t0)
s, k = slv.last_solve_stats()
it = slv.iterations_number
tt = torch.tensor([best], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"case": f"block4x4 {nx}x{ny}x{nzl}/GPU AMG+DILU {outer} {mode}", "n_gpus": world, "block_rows_global": ng, "iters": it, "solve_s_wall_max": float(tt.item()),
                      "solve_s_events_rank0": s, "iters_per_s_global": it / float(tt.item()), "value_weak": it / float(tt.item()) * world, "status": slv.status,
                      "launches_rank0": k, "levels": slv.num_levels(), "colors_L0": slv.level_coloring(0)[0], "setup_s": tsetup, "gen_s": tgen}), flush=True)
for o in (slv, x, b, A, rsc, cfg):
    o.destroy()
capi.finalize()
if world > 1:
    dist.destroy_process_group()
