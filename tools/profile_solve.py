"""Profile harness: setup + warm-up solve outside the profiled range, then ONE solve of a few iterations between
cudaProfilerStart/Stop (use with ncu --profile-from-start off)."""
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfgd = json.loads((ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json").read_text())
cfgd["solver"]["max_iters"] = iters
cfgd["solver"]["tolerance"] = 1e-30
capi.initialize()
capi.register_print_callback(None)
cfg = capi.Config(cfgd)
rsc = capi.Resources(cfg)
A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
A.generate_poisson7(b, x, nx, nx, nx)
n, _, _ = A.get_size()
slv = capi.Solver(rsc, cfg)
slv.setup(A)
x.set_zero(n)
slv.solve(b, x, zero_initial_guess=True)
rt = ctypes.CDLL("libcudart.so")
x.set_zero(n)
rt.cudaProfilerStart()
slv.solve(b, x, zero_initial_guess=True)
rt.cudaProfilerStop()
s, k = slv.last_solve_stats()
print(f"profiled solve: n={nx}^3 iters={slv.iterations_number} levels={slv.num_levels()} {s*1e3:.3f} ms {k} launches")
for l in range(slv.num_levels()):
    print(l, slv.level_info(l))
