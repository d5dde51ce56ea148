import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
capi.initialize(); capi.register_print_callback(None)
cfg = capi.Config(file=str(ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
rsc = capi.Resources(cfg, device=lr, comm=capi.AMGXB200_comm(rank, world, idt.cpu().numpy().tobytes()))
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
A.generate_poisson7(b, x, nx, nx, nx, 1, 1, world)
n = A.get_size()[0]
slv = capi.Solver(rsc, cfg)
slv.setup(A)
x.set_zero(n)
slv.solve(b, x, zero_initial_guess=True)
lv = [slv.level_info(l)["n"] for l in range(slv.num_levels())]
hist = slv.residual_history()
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: iters {slv.iterations_number} {slv.status} levels {lv} hist[:4] {[float(h) for h in hist[:4]]}", flush=True)
capi.finalize(); dist.destroy_process_group()
