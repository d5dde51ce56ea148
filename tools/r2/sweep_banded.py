"""Fine-level kernel times on the SuiteSparse-shaped banded matrix (gallery.random_banded defaults: 4 M rows) through AMGXB200_bench_kernel;
also the target of the ncu capture of the window kernel (3 warm-up + `reps` launches per kernel kind: SpMV, fused Jacobi, SpMV + dot)."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
from amgx_b200 import capi, gallery

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rp, ci, va = gallery.random_banded(n)
capi.initialize()
capi.register_print_callback(None)
cfg = capi.Config("config_version=2, solver(main)=NOSOLVER")
rsc = capi.Resources(cfg)
A = capi.Matrix(rsc).upload(rp, ci, va)
nnz = int(rp[-1])
byt = nnz * 12 + n * 4
peak = 6575.1
out = {"n": n, "nnz": nnz, "plan": A.kernel_info(), "knobs": {k: os.environ[k] for k in os.environ if k.startswith("AMGXB_")}}
for kind, name, extra in ((0, "spmv", 0), (1, "jacobi", 4 * n * 8), (2, "spmv_dot", 0)):
    ms = A.bench_kernel(kind, warmup=1, reps=reps)
    out[name] = {"ms": round(ms, 4), "frac": round((byt + extra) / ms / 1e6 / peak, 4)}
print(json.dumps(out), flush=True)
for o in (A, rsc, cfg):
    o.destroy()
capi.finalize()
