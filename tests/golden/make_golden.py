"""Generate golden vectors by running the UNMODIFIED reference on a B200 (gpurun):

    gpurun -- 'python tests/golden/make_golden.py'          # everything
    gpurun -- 'python tests/golden/make_golden.py r2'       # only the round-2 cases (tests/golden/cases_round2.py)

For every case it writes the system, runs oracle/_ref/ref_dump (reference's own C API + a dump of its
internal hierarchy) with a JSON config in the reference's format, and stores a compressed fixture in
gpurun_out/golden/<case>.npz; copy those to tests/golden/ and commit them.  Nothing here touches
/root/reference (it does not exist on the GPU box): the reference is the prebuilt oracle/_ref/."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from amgx_b200 import gallery  # noqa: E402
from tests.golden.refdump_io import read_dump, write_system  # noqa: E402

OUT = ROOT / "gpurun_out" / "golden"
REF = ROOT / "oracle" / "_ref" / "ref_dump"


def cfg_pcg_agg(tol=1e-10, max_iters=60, pre=0, post=3, omega=0.8, coarsest=2):
    return {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "PCG", "max_iters": max_iters, "monitor_residual": 1, "store_res_history": 1,
        "convergence": "RELATIVE_INI", "tolerance": tol, "norm": "L2", "print_solve_stats": 0, "obtain_timings": 0,
        "preconditioner": {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V",
                           "max_levels": 50, "presweeps": pre, "postsweeps": post, "coarsest_sweeps": coarsest, "coarse_solver": "NOSOLVER",
                           "max_iters": 1, "monitor_residual": 0, "print_grid_stats": 1,
                           "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": omega, "monitor_residual": 0}}}}


def cfg_pcg_jacobi(tol=1e-8, max_iters=40):
    return {"config_version": 2, "solver": {
        "scope": "main", "solver": "PCG", "max_iters": max_iters, "monitor_residual": 1, "store_res_history": 1,
        "convergence": "RELATIVE_INI", "tolerance": tol, "norm": "L2",
        "preconditioner": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}}}


def cfg_amg_agg_standalone(tol=1e-8, max_iters=40, pre=1, post=1, norm="L1"):
    return {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
        "presweeps": pre, "postsweeps": post, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": max_iters,
        "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": tol, "norm": norm,
        "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}}}


def cfg_fgmres_agg(tol=1e-10, max_iters=60, restart=10, precond=True):
    c = {"config_version": 2, "determinism_flag": 1, "solver": {
        "scope": "main", "solver": "FGMRES", "max_iters": max_iters, "gmres_n_restart": restart, "monitor_residual": 1, "store_res_history": 1,
        "convergence": "RELATIVE_INI", "tolerance": tol, "norm": "L2",
        "preconditioner": {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V",
                           "max_levels": 50, "presweeps": 1, "postsweeps": 2, "coarse_solver": "NOSOLVER",
                           "max_iters": 1, "monitor_residual": 0,
                           "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.75, "monitor_residual": 0}}}}
    if not precond:
        c["solver"]["preconditioner"] = {"scope": "nop", "solver": "NOSOLVER"}
    return c


def cfg_amg_dilu(tol=1e-8, max_iters=40, norm="L1", determinism=0):
    return {"config_version": 2, "determinism_flag": determinism, "solver": {
        "scope": "main", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
        "matrix_coloring_scheme": "MIN_MAX", "max_uncolored_percentage": 0.15, "smoother": "MULTICOLOR_DILU", "relaxation_factor": 0.9,
        "presweeps": 1, "postsweeps": 1, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": max_iters,
        "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": tol, "norm": norm}}


def cfg_fgmres_classical(tol=1e-10, max_iters=60, restart=20, interpolator="D2", aggressive_levels=1, max_elements=4, strength_threshold=0.25,
                         max_row_sum=0.9, pre=2, post=2):
    """src/configs/FGMRES_CLASSICAL_AGGRESSIVE_PMIS.json (BASELINE config 3) with a tighter tolerance"""
    amg = {"scope": "amg_solver", "solver": "AMG", "algorithm": "CLASSICAL", "selector": "PMIS", "interpolator": interpolator,
           "aggressive_levels": aggressive_levels, "interp_max_elements": max_elements, "max_row_sum": max_row_sum,
           "strength_threshold": strength_threshold, "cycle": "V", "max_levels": 50, "min_coarse_rows": 2, "presweeps": pre, "postsweeps": post,
           "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 1, "monitor_residual": 0, "print_grid_stats": 1,
           "smoother": {"scope": "jacobi", "solver": "JACOBI_L1", "relaxation_factor": 1, "monitor_residual": 0}}
    return {"config_version": 2, "solver": {
        "scope": "main", "solver": "FGMRES", "max_iters": max_iters, "gmres_n_restart": restart, "monitor_residual": 1, "store_res_history": 1,
        "convergence": "RELATIVE_INI", "tolerance": tol, "norm": "L2", "preconditioner": amg}}


def cfg_pcg_agg_block(tol=1e-8, max_iters=60):
    c = cfg_pcg_agg(tol=tol, max_iters=max_iters, pre=1, post=1, omega=0.9)
    return c


def cases():
    ex = np.load(ROOT / "tests" / "golden" / "example_matrix_12x12.npz")
    yield "example12_pcg_jacobi", (ex["row_ptr"], ex["col_idx"], ex["values"]), cfg_pcg_jacobi(), None
    for n in (6, 10, 16):
        yield f"poisson{n}_pcg_agg_jacobi", gallery.poisson7pt(n), cfg_pcg_agg(), None
    yield "poisson12x10x7_pcg_agg_jacobi", gallery.poisson7pt(12, 10, 7), cfg_pcg_agg(), None
    yield "poisson10_sorted_pcg_agg_jacobi", gallery.poisson7pt_sorted(10), cfg_pcg_agg(), None
    yield "poisson12_pcg_agg_jacobi_pre2", gallery.poisson7pt(12), cfg_pcg_agg(pre=2, post=2, omega=0.7), None
    yield "poisson12_amg_standalone_L1", gallery.poisson7pt(12), cfg_amg_agg_standalone(), None
    yield "banded3000_pcg_agg_jacobi", gallery.random_banded(3000, sigma=40.0), cfg_pcg_agg(max_iters=80), None
    yield "poisson8_pcg_jacobi", gallery.poisson7pt(8), cfg_pcg_jacobi(tol=1e-10, max_iters=80), None
    yield "poisson12_amg_dilu", gallery.poisson7pt(12), cfg_amg_dilu(), None
    yield "poisson9_amg_dilu_det", gallery.poisson7pt(9), cfg_amg_dilu(determinism=1, norm="L2"), None
    yield "block4_6x5x4_pcg_agg_bjacobi", gallery.block_elasticity(6, 5, 4), cfg_pcg_agg_block(), (4, "dDDI")
    yield "block4_6x5x4_amg_dilu", gallery.block_elasticity(6, 5, 4), cfg_amg_dilu(tol=1e-8, max_iters=30, norm="L2"), (4, "dDDI")
    yield "block4_6x5x4_amg_dilu_dDFI", gallery.block_elasticity(6, 5, 4), cfg_amg_dilu(tol=1e-5, max_iters=30, norm="L2"), (4, "dDFI")
    yield "poisson10_fgmres_agg_jacobi", gallery.poisson7pt(10), cfg_fgmres_agg(restart=5), None
    yield "poisson8_fgmres_noprec", gallery.poisson7pt(8), cfg_fgmres_agg(restart=12, max_iters=70, tol=1e-8, precond=False), None
    yield "banded3000_fgmres_agg_jacobi", gallery.random_banded(3000, sigma=40.0), cfg_fgmres_agg(restart=8, max_iters=40), None
    # classical AMG (config 3 family); the third entry asks ref_dump for the level-0 stage dump (strength, C/F maps, P before truncation)
    yield "poisson12_fgmres_classical_aggr", gallery.poisson7pt(12), cfg_fgmres_classical(), (1, "dDDI", "0.25,0.9,4")
    yield "poisson16x12x9_fgmres_classical_aggr", gallery.poisson7pt(16, 12, 9), cfg_fgmres_classical(), (1, "dDDI", "0.25,0.9,4")
    yield "poisson12_sorted_fgmres_classical_d2", gallery.poisson7pt_sorted(12), cfg_fgmres_classical(aggressive_levels=0, max_elements=-1), (1, "dDDI", "0.25,0.9,-1")
    yield "banded3000_fgmres_classical_d2_trunc", gallery.random_banded(3000, sigma=40.0), cfg_fgmres_classical(aggressive_levels=0, max_iters=40), (1, "dDDI", "0.25,0.9,4")


def all_cases():
    yield from cases()
    # components written after round 1's GPU minutes were spent: same definitions the oracle / engine tests use
    from tests.golden.cases_round2 import cases as cases_r2
    for name, mat, cfg in cases_r2():
        yield "r2_" + name, mat, cfg, None


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    only = sys.argv[1:] or None
    if only == ["r2"]:
        only = ["r2_" + n for n, _, _ in __import__("tests.golden.cases_round2", fromlist=["cases"]).cases()]
    for name, (rp, ci, va), cfg, extra in all_cases():
        if only and name not in only:
            continue
        bs, mode = (extra[0], extra[1]) if extra else (1, "dDDI")
        env = dict(os.environ)
        if extra and len(extra) > 2:
            env["REFDUMP_CLASSICAL"] = extra[2]
        n = rp.shape[0] - 1
        rhs = np.ones(n * bs)
        sysf, cfgf, outf = OUT / f"{name}.sys", OUT / f"{name}.json", OUT / f"{name}.bin"
        write_system(sysf, rp, ci, va, rhs, block=(bs, bs))
        cfgf.write_text(json.dumps(cfg, indent=1))
        try:
            r = subprocess.run([str(REF), str(sysf), str(cfgf), str(outf), mode], capture_output=True, text=True, env=env, timeout=180)
        except subprocess.TimeoutExpired:
            print(f"[{name}] ref_dump TIMED OUT after 180 s")
            continue
        if r.returncode != 0:
            print(f"[{name}] ref_dump FAILED rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        d = read_dump(outf)
        d["config_json"] = json.dumps(cfg)
        d["sys_row_ptr"], d["sys_col_idx"], d["sys_values"], d["sys_rhs"] = rp, ci, va, rhs
        d["sys_block"], d["sys_mode"] = np.array([bs]), mode
        np.savez_compressed(OUT / f"{name}.npz", **{k: v for k, v in d.items()})
        print(f"[{name}] n={n} status={d['status'][0]} iters={d['iterations'][0]} levels={d.get('num_levels', [0])[0]} "
              f"res0={d['res_history'][0]:.6e} resN={d['res_history'][-1]:.6e}")
        for f in (sysf, outf):
            f.unlink()


if __name__ == "__main__":
    main()
