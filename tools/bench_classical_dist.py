"""Config 3 (FGMRES + classical AMG) on a row-partitioned nx * nx * nz Poisson problem, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29731 tools/bench_classical_dist.py NX [NZ_TOTAL]
Strong scaling: the GLOBAL grid is fixed (nx * nx * nz_total, z-slabs of nz_total / N planes per rank), so the lines for N = 1, 2, 4, 8 are the same
problem and must report the same iteration count (the hierarchy is the single-GPU hierarchy of the global matrix).
Times: CUDA events inside the engine (last_solve_stats), max over ranks.  NOT yet run on a device (written after round 1's GPU minutes)."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi  # noqa: E402
from tests.golden.make_golden import cfg_fgmres_classical  # noqa: E402


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nz_total = int(sys.argv[2]) if len(sys.argv) > 2 else nx
    assert nz_total % world == 0, "nz_total must be a multiple of the number of ranks"
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    capi.initialize()
    capi.register_print_callback(None)
    cfgd = cfg_fgmres_classical(tol=1e-6, max_iters=100, restart=20)
    cfgd["solver"]["preconditioner"]["print_grid_stats"] = 0
    cfg = capi.Config(cfgd)
    comm = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.AMGXB200_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
    rsc = capi.Resources(cfg, device=lr, comm=comm) if world > 1 else capi.Resources(cfg)
    A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
    A.generate_poisson7(b, x, nx, nx, nz_total // world, 1, 1, world)
    n = A.get_size()[0]
    slv = capi.Solver(rsc, cfg)
    t = time.time()
    slv.setup(A)
    torch.cuda.synchronize()
    ts = time.time() - t
    if world > 1:
        b.bind(A)
        x.bind(A)
    best = 1e30
    for _ in range(3):
        x.set_zero(n)
        if world > 1:
            dist.barrier()
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        if world > 1:
            tt = torch.tensor([s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            s = float(tt.item())
        best = min(best, s)
    it = slv.iterations_number
    hist = slv.residual_history()
    if rank == 0:
        print(json.dumps({"engine": "ours", "n_gpus": world, "grid": [nx, nx, nz_total], "rows": nx * nx * nz_total, "setup_s": ts, "solve_s": best, "iters": it,
                          "iters_per_s": it / best, "launches": k, "num_levels": slv.num_levels(), "final_rel": float(hist[-1] / hist[0]), "status": slv.status}), flush=True)
    for o in (slv, x, b, A, rsc, cfg):
        o.destroy()
    capi.finalize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
