"""Ad-hoc kernel timing on the GPU box: SpMV / fused Jacobi / SpMV+dot GB/s on generated Poisson matrices."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi  # noqa: E402

capi.initialize()
cfg = capi.Config("config_version=2, solver(main)=NOSOLVER")
rsc = capi.Resources(cfg)
sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
for nx in sizes:
    A = capi.Matrix(rsc)
    t = time.time()
    A.generate_poisson7(None, None, nx, nx, nx)
    n, _, _ = A.get_size()
    nnz = A.get_nnz()
    gen = time.time() - t
    byt = nnz * 12 + n * 4
    for kind, name in ((0, "spmv"), (1, "jacobi"), (2, "spmv_dot")):
        for flush in (False, True):
            ms = A.bench_kernel(kind, warmup=3, reps=20, flush_l2=flush)
            print(f"n={nx}^3 rows={n} nnz={nnz} {name:9s} flush={int(flush)} {ms:8.4f} ms  {byt/ms/1e6:8.1f} GB/s (north-star bytes)  gen {gen:.2f}s", flush=True)
    A.destroy()
