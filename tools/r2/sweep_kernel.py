"""One process = one knob setting (the knobs are read once per process): fine-level kernel times on n^3 Poisson through AMGXB200_bench_kernel.
usage: [AMGXB_TILE_STAGES=..] [AMGXB_TILE_CTAS=..] [AMGXB_TILE_UNROLL=..] [AMGXB_COLENC=..] python tools/r2/sweep_kernel.py 256 [solve]"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
from amgx_b200 import capi

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
do_solve = len(sys.argv) > 2 and sys.argv[2] == "solve"
capi.initialize()
capi.register_print_callback(None)
cfg = capi.Config(file=str(ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"))
rsc = capi.Resources(cfg)
A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
A.generate_poisson7(b, x, nx, nx, nx, 1, 1, 1)
n, _, _ = A.get_size()
nnz = A.get_nnz()
byt = nnz * 12 + n * 4
out = {"nx": nx, "knobs": {k: os.environ[k] for k in os.environ if k.startswith("AMGXB_")}}
peak = 6575.1
for kind, name, extra in ((0, "spmv", 0), (1, "jacobi", 4 * n * 8), (2, "spmv_dot", 0)):
    ms = A.bench_kernel(kind, warmup=3, reps=20)
    out[name] = {"ms": round(ms, 4), "frac": round((byt + extra) / ms / 1e6 / peak, 4)}
if do_solve:
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    tot_s = tot_it = 0
    for i in range(4):
        x.set_zero(n)
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        if i:
            tot_s += s
            tot_it += slv.iterations_number
    out["solve"] = {"iters": slv.iterations_number, "its_per_s": round(tot_it / tot_s, 1), "status": slv.status}
    slv.destroy()
print(json.dumps(out), flush=True)
for o in (x, b, A, rsc, cfg):
    o.destroy()
capi.finalize()
