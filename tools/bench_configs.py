"""Extra measurements for the other BASELINE configs (not bench.py lines): 512^3 aggregation PCG on one GPU, and the
block-4x4 multicolour-DILU config in dDFI / dDDI."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi, gallery  # noqa: E402

capi.initialize()
capi.register_print_callback(None)
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def poisson(nx):
    cfg = capi.Config(file=str(ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"))
    rsc = capi.Resources(cfg)
    A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
    t = time.time()
    A.generate_poisson7(b, x, nx, nx, nx)
    n = A.get_size()[0]
    nnz = A.get_nnz()
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    ts = time.time() - t
    for _ in range(2):
        x.set_zero(n)
        slv.solve(b, x, zero_initial_guess=True)
    s, k = slv.last_solve_stats()
    it = slv.iterations_number
    byt = nnz * 12 + n * 4
    ms = A.bench_kernel(0, 3, 10)
    msj = A.bench_kernel(1, 3, 10)
    print(json.dumps({"case": f"poisson{nx}^3 PCG+AGG+JACOBI", "rows": n, "nnz": nnz, "iters": it, "solve_s": s, "iters_per_s": it / s, "status": slv.status,
                      "setup_s": ts, "levels": slv.num_levels(), "spmv_ms": ms, "spmv_GBs_northstar": byt / ms / 1e6, "spmv_frac_of_6575": byt / ms / 1e6 / 6575.1,
                      "jacobi_ms": msj, "jacobi_GBs": (byt + 32 * n) / msj / 1e6, "jacobi_frac": (byt + 32 * n) / msj / 1e6 / 6575.1}), flush=True)
    for o in (slv, x, b, A, rsc, cfg):
        o.destroy()


def block(mode, nx, ny, nz, scheme="MIN_MAX"):
    cfgd = {"config_version": 2, "solver": {
        "scope": "main", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
        "matrix_coloring_scheme": scheme, "max_uncolored_percentage": 0.15, "smoother": "MULTICOLOR_DILU", "relaxation_factor": 0.9,
        "presweeps": 1, "postsweeps": 1, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 100,
        "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": 1e-6, "norm": "L2"}}
    cfg = capi.Config(cfgd)
    rsc = capi.Resources(cfg)
    rp, ci, va = gallery.block_elasticity(nx, ny, nz, dtype=np.float32 if mode == "dDFI" else np.float64)
    n = rp.shape[0] - 1
    A = capi.Matrix(rsc, mode).upload(rp, ci, va, block_dims=(4, 4))
    b = capi.Vector(rsc, mode).upload(np.ones(n * 4), block_dim=4)
    x = capi.Vector(rsc, mode).set_zero(n, 4)
    slv = capi.Solver(rsc, cfg, mode)
    t = time.time()
    slv.setup(A)
    ts = time.time() - t
    for _ in range(2):
        x.set_zero(n, 4)
        slv.solve(b, x, zero_initial_guess=True)
    s, k = slv.last_solve_stats()
    it = slv.iterations_number
    nnzb = ci.shape[0]
    byt = nnzb * (16 * (4 if mode == "dDFI" else 8) + 4) + n * 4
    ms = A.bench_kernel(0, 3, 10)
    hist = slv.residual_history()
    print(json.dumps({"case": f"block4x4 {nx}x{ny}x{nz} AMG+DILU {mode} {scheme}", "block_rows": n, "nnz_blocks": nnzb, "iters": it, "solve_s": s, "iters_per_s": it / s,
                      "status": slv.status, "setup_s": ts, "levels": slv.num_levels(), "colors_L0": slv.level_coloring(0)[0], "launches": k,
                      "spmv_ms": ms, "spmv_GBs_northstar": byt / ms / 1e6, "spmv_frac_of_6575": byt / ms / 1e6 / 6575.1, "final_rel": hist[-1] / hist[0]}), flush=True)
    for o in (slv, x, b, A, rsc, cfg):
        o.destroy()


if what in ("all", "p512"):
    poisson(512)
if what in ("all", "block"):
    block("dDFI", 160, 160, 160)
    block("dDDI", 160, 160, 160)
if what == "block_pg":      # PARALLEL_GREEDY colouring: about half the colours of MIN_MAX on a 7-point block stencil = half the launches per DILU sweep
    block("dDFI", 160, 160, 160, "MIN_MAX")
    block("dDFI", 160, 160, 160, "PARALLEL_GREEDY")
