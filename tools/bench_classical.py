"""Config 3 (FGMRES + classical AMG, aggressive PMIS / MULTIPASS level 0, D2 below, trunc 4, JACOBI_L1) on nx^3 Poisson:
this engine vs the reference GPU build (oracle/_ref/ref_dump), same JSON config."""
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi  # noqa: E402
from tests.golden.make_golden import cfg_fgmres_classical  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
capi.initialize()
capi.register_print_callback(None)
for nx in sizes:
    cfgd = cfg_fgmres_classical(tol=1e-6, max_iters=100, restart=20)
    cfgd["solver"]["preconditioner"]["print_grid_stats"] = 0
    cfgf = OUT / f"cla_{nx}.json"
    cfgf.write_text(json.dumps(cfgd))
    cfg = capi.Config(cfgd)
    rsc = capi.Resources(cfg)
    A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
    A.generate_poisson7(b, x, nx, nx, nx)
    n = A.get_size()[0]
    slv = capi.Solver(rsc, cfg)
    t = time.time()
    slv.setup(A)
    ts = time.time() - t
    best = 1e30
    for _ in range(3):
        x.set_zero(n)
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        best = min(best, s)
    it = slv.iterations_number
    lv = [(slv.level_info(l)["n"], slv.level_info(l)["nnz"]) for l in range(slv.num_levels())]
    hist = slv.residual_history()
    print(json.dumps({"engine": "ours", "nx": nx, "rows": n, "setup_s": ts, "solve_s": best, "iters": it, "iters_per_s": it / best, "launches": k,
                      "levels": lv[:8], "num_levels": len(lv), "final_rel": hist[-1] / hist[0], "status": slv.status}), flush=True)
    for o in (slv, x, b, A, rsc, cfg):
        o.destroy()
    outf = OUT / f"cla_ref_{nx}.bin"
    t = time.time()
    r = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_dump"), f"poisson:{nx}", str(cfgf), str(outf), "dDDI", "3"], capture_output=True, text=True)
    if r.returncode != 0:
        print(json.dumps({"engine": "reference", "nx": nx, "error": (r.stdout[-300:] + r.stderr[-300:])}), flush=True)
        continue
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from refdump_io import read_dump
    d = read_dump(outf)
    it_r = int(d["iterations"][0])
    print(json.dumps({"engine": "reference", "nx": nx, "setup_s": float(d["times"][0]), "solve_s": float(d["times"][1]), "iters": it_r,
                      "iters_per_s": it_r / float(d["times"][1]), "final_rel": float(d["res_history"][-1] / d["res_history"][0]),
                      "wall_s": time.time() - t}), flush=True)
    outf.unlink()
