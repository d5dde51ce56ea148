"""Reference GPU build on the config-5 style problem (4x4 blocks, AMG + MULTICOLOR_DILU), same JSON as tools/bench_block_dist.py."""
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from amgx_b200 import gallery  # noqa: E402
from refdump_io import read_dump, write_system  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "dDFI"
nx, ny, nz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (128, 128, 80)
outer = sys.argv[5] if len(sys.argv) > 5 else "PCG"
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
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
rp, ci, va = gallery.block_elasticity(nx, ny, nz)
n = rp.shape[0] - 1
sysf, cfgf, outf = out / "blk.sys", out / "blk.json", out / "blk.bin"
write_system(sysf, rp, ci, va, np.ones(n * 4), block=(4, 4))
cfgf.write_text(json.dumps(cfgd))
t = time.time()
env = dict(REFDUMP_NO_LEVELS="1")
import os
env = dict(os.environ, REFDUMP_NO_LEVELS="1")
r = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_dump"), str(sysf), str(cfgf), str(outf), mode, "3"], capture_output=True, text=True, env=env)
if r.returncode != 0:
    print(json.dumps({"engine": "reference", "error": r.stdout[-400:] + r.stderr[-400:]}))
else:
    d = read_dump(outf)
    it = int(d["iterations"][0])
    print(json.dumps({"engine": "reference", "case": f"block4x4 {nx}x{ny}x{nz} AMG+DILU {outer} {mode}", "block_rows": n, "iters": it, "setup_s": float(d["times"][0]),
                      "solve_s": float(d["times"][1]), "iters_per_s": it / float(d["times"][1]), "status": int(d["status"][0]), "wall_s": time.time() - t}))
for f in (sysf, outf):
    if f.exists():
        f.unlink()
