"""2+ GPUs (torchrun): latency of one distributed matrix apply (halo exchange + SpMV) and of one PCG+Jacobi iteration on SMALL row-partitioned
problems -- what the coarse levels of the hierarchy pay.  Compare AMGXB_P2P=1 (peer-memory kernels) with AMGXB_P2P=0 (NCCL send/recv)."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist
from amgx_b200 import capi

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
capi.initialize()
capi.register_print_callback(None)
cfgd = {"config_version": 2, "determinism_flag": 1, "solver": {"scope": "main", "solver": "PCG", "max_iters": 200, "monitor_residual": 1, "convergence": "RELATIVE_INI",
        "tolerance": 1e-30, "norm": "L2", "preconditioner": {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}}}
cfg = capi.Config(cfgd)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
rsc = capi.Resources(cfg, device=lr, comm=capi.AMGXB200_comm(rank, world, bytes(idt.cpu().numpy().tobytes())))
out = {"world": world, "p2p": os.environ.get("AMGXB_P2P", "1")}
for g in (16, 48, 128):
    A, b, x, y = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc), capi.Vector(rsc)
    A.generate_poisson7(b, x, g, g, g, 1, 1, world)
    n = g * g * g
    for v in (b, x, y):
        v.bind(A)
    y.set_zero(n)
    for _ in range(20):
        A.multiply(x, y)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        A.multiply(x, y)
    torch.cuda.synchronize()
    mult_us = (time.perf_counter() - t0) / 300 * 1e6
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    for _ in range(2):
        x.set_zero(n)
        slv.solve(b, x, zero_initial_guess=True)
    s, k = slv.last_solve_stats()
    out[f"g{g}"] = {"multiply_us_host_timed": round(mult_us, 1), "pcg_jacobi_iteration_us": round(s / slv.iterations_number * 1e6, 1), "launches_per_iteration": round(k / slv.iterations_number, 1)}
    for o in (slv, y, x, b, A):
        o.destroy()
if rank == 0:
    print(json.dumps(out), flush=True)
rsc.destroy(); cfg.destroy(); capi.finalize(); dist.destroy_process_group()
