#!/usr/bin/env python
"""bench.py -- solve-phase benchmark of the B200 AMG engine (contract: see README / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload poisson|classical|banded|block] [--grid NX] [--strong]

Default workload (BASELINE.json configs[1]): 3-D 7-point Poisson 256^3, fp64, aggregation AMG (SIZE_2) V-cycle preconditioned CG,
BLOCK_JACOBI(0.8) 0+3 sweeps -- the reference's PCG_AGGREGATION_JACOBI.json.  A "step" is one AMGX_solver_solve (zero initial guess,
b = 1, RELATIVE_INI 1e-6, max 100 iterations); metric = outer Krylov iterations (one V-cycle each) per second, whole job.
  value : solves with b / x already resident in HBM, timed by CUDA events the library records on its solve stream
          (AMGXB200_solver_get_last_solve_stats), max over ranks.
  e2e   : the same solve through the C-ABI with HOST buffers: AMGX_vector_upload(rhs) from pinned host memory + AMGX_vector_set_zero
          + AMGX_solver_solve + AMGX_vector_download(sol), wall clock around the synchronous calls.
  roofline : the DOMINANT kernel of the iteration = the fused Jacobi sweep on the fine level (csr_tile_kernel<EPI_JACOBI>, ~70 % of an
          iteration: profiles/r02_launches_solve_256.md); algorithmic bytes nnz*(8+4) + rows*4 + 4*rows*8 per launch / CUDA-event time
          per launch, against the measured HBM copy peak (MEASURED_PEAKS.json).  roofline.spmv carries the plain fine-level SpMV
          (north-star bytes nnz*12 + rows*4), roofline.iteration the whole outer iteration.
  reference_gpu : the UNMODIFIED reference (oracle/_ref, its own sm_100 GPU build) on the same matrix and configuration, timed in
          this run by its own harness (oracle/ref_build/ref_dump.cu, cudaEvents around AMGX_solver_solve).  Context, not the target.
  cpu_baseline / --impl reference : the CPU oracle port of the reference algorithm (the reference's own host path has no SIZE_2
          aggregation: src/aggregation/selectors/size2_selector.cu:630-643 throws) at the REAL problem size, on an explicit
          thread count, bounded in iterations.
N > 1 (one process per GPU, row partition in z-slabs, peer-memory / NCCL halo exchange):
  default  weak scaling: every rank owns an NX^3 box, global grid NX x NX x (NX*N); value = global iterations/s x N (see
           config.value_definition), config.global_iterations_per_sec is the raw rate.
  --strong the global grid stays NX^3 (BASELINE configs[3] with --grid 512): value = global iterations/s, scaling "strong".
  parity   object: the same distributed code path on a reduced global problem against a single-rank solve of the assembled matrix.
  strong_512 object (default weak run, N divides 512): BASELINE configs[3] as stated -- the 512^3 grid row-partitioned over the N ranks,
           same solver, device-timed like `value`; its N=1 counterpart is other_workloads.poisson512 of the N=1 line.
  block_weak object (same runs): BASELINE configs[4]-style weak scaling, 160^3 4x4 block rows per GPU, AMG + MULTICOLOR_DILU, dDFI;
           N=1 counterpart other_workloads.block160_dDFI.
Default N=1 line only: other_workloads = the other BASELINE workloads (512^3 Poisson = the north star's target size, FGMRES + classical AMG at
  512^3 = configs[2], the 4 M-row SuiteSparse-shaped matrix, the 4x4 block configuration), each measured by THIS script in a child process after the main line's
  numbers are final (`python bench.py --workload ... --no-extras`), under a common time budget; --no-extras skips them.
  reference_host_path = the reference's own host implementation (mode hDDI, PCG + BLOCK_JACOBI: the solver family it has on the CPU),
  timed on the host cores beside this engine on the same configuration.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIG = ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"
REF_SYSTEM = "/tmp/amgxb_ref_system.bin"
METRIC = "solve_phase_vcycle_iterations_per_sec"
UNIT = "iterations/s"

BLOCK_CFG = {"config_version": 2, "solver": {
    "scope": "main", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
    "matrix_coloring_scheme": "MIN_MAX", "max_uncolored_percentage": 0.15, "smoother": "MULTICOLOR_DILU", "relaxation_factor": 0.9,
    "presweeps": 1, "postsweeps": 1, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 100,
    "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": 1e-6, "norm": "L2"}}


# SuiteSparse-shaped CSR is not symmetric: FGMRES outside, the same aggregation-AMG V-cycle (BLOCK_JACOBI 0.8, 0 + 3 sweeps) inside
BANDED_CFG = {"config_version": 2, "determinism_flag": 1, "solver": {
    "scope": "main", "solver": "FGMRES", "max_iters": 100, "gmres_n_restart": 20, "monitor_residual": 1, "store_res_history": 1,
    "convergence": "RELATIVE_INI", "tolerance": 1e-6, "norm": "L2",
    "preconditioner": {"scope": "amg", "solver": "AMG", "algorithm": "AGGREGATION", "selector": "SIZE_2", "cycle": "V", "max_levels": 50,
                       "presweeps": 0, "postsweeps": 3, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 1, "monitor_residual": 0,
                       "smoother": {"scope": "jacobi", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "monitor_residual": 0}}}}


# BASELINE configs[2]: the reference's FGMRES_CLASSICAL_AGGRESSIVE_PMIS.json (PMIS, aggressive level 0 with MULTIPASS, D2 below, interp_max_elements 4,
# JACOBI_L1 2 + 2 sweeps) with gmres_n_restart 20 (100 Krylov vectors of 512^3 rows would not fit beside the hierarchy)
CLASSICAL_CFG = {"config_version": 2, "solver": {
    "scope": "main", "solver": "FGMRES", "max_iters": 100, "gmres_n_restart": 20, "monitor_residual": 1, "store_res_history": 1,
    "convergence": "RELATIVE_INI", "tolerance": 1e-6, "norm": "L2",
    "preconditioner": {"scope": "amg_solver", "solver": "AMG", "algorithm": "CLASSICAL", "selector": "PMIS", "interpolator": "D2", "aggressive_levels": 1,
                       "interp_max_elements": 4, "max_row_sum": 0.9, "strength_threshold": 0.25, "cycle": "V", "max_levels": 50, "min_coarse_rows": 2,
                       "presweeps": 2, "postsweeps": 2, "coarsest_sweeps": 2, "coarse_solver": "NOSOLVER", "max_iters": 1, "monitor_residual": 0,
                       "print_grid_stats": 0,
                       "smoother": {"scope": "jacobi", "solver": "JACOBI_L1", "relaxation_factor": 1, "monitor_residual": 0}}}}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores, at the real size of the workload, bounded in iterations
# ------------------------------------------------------------------------------------------------------------------------
def host_threads():
    """Explicit thread count of the CPU arm: the cores this process may run on, capped at 32 (these memory-bound loops lose beyond
    that).  torchrun's OMP_NUM_THREADS=1 is deliberately ignored: the baseline gets the cores of the box whatever launched it."""
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    return max(1, min(avail, 32)), avail


def oracle_baseline(nx, ny, nz, iters):
    """PCG + aggregation-AMG of the oracle on the nx x ny x nz Poisson problem: `iters` iterations, timed; setup untimed."""
    from amgx_b200 import gallery
    from oracle import oracle as orc
    cores, avail = host_threads()
    orc.set_num_threads(cores)
    rp, ci, va = gallery.poisson7pt(nx, ny, nz)
    n = rp.shape[0] - 1
    t0 = time.time()
    amg = orc.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.8)
    t_setup = time.time() - t0
    orc.pcg(rp, ci, va, np.ones(n), amg=amg, tol=1e-30, max_iters=1)          # touch every level once (page faults, thread pool)
    t0 = time.time()
    _, it, hist, _ = orc.pcg(rp, ci, va, np.ones(n), amg=amg, tol=1e-30, max_iters=iters)
    dt = time.time() - t0
    return {"value": it / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle PCG + aggregation-AMG on the full 7-pt Poisson {nx}x{ny}x{nz} ({n} rows, {amg.num_levels()} levels): {it} iterations in {dt:.2f} s on "
                      f"{cores} OpenMP threads ({avail} host cores visible); setup {t_setup:.1f} s untimed; no extrapolation"}


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference has no CPU SIZE_2), same config."""
    if rank != 0:
        return
    nx = args.n
    nz = nx if (args.strong or world == 1) else nx * world
    note = None
    if world > 2 and not args.strong:
        # the 8x problem needs > 60 GB of host arrays in numpy: bound the sample to two slabs and say so
        nz = nx * 2
        note = f"sample = 2 of the {world} slabs (host memory bound); iterations/s scaled by rows ratio 2/{world}"
    iters = 5 if nx <= 256 else 2
    vals, base = [], None
    for _ in range(max(1, min(args.steps, 2))):
        base = oracle_baseline(nx, nx, nz, iters)
        vals.append(base["value"])
    v = float(np.median(vals))
    if note:
        v *= 2.0 / world
        base["sample"] += "; " + note
    if not args.strong:
        v *= world          # same normalisation as the engine's line: N=1-sized sub-problems advanced per second
    base["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": None, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world),
           "cpu_baseline": base,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference has no CPU implementation of SIZE_2 aggregation (size2_selector.cu:630-643 throws); this is the oracle port"}
    print(json.dumps(out), flush=True)


def workload_config(args, world):
    nx = args.n
    if args.workload == "poisson":
        if args.strong:
            w = f"7-pt Poisson {nx}^3 fp64 row-partitioned over {world} GPU(s), PCG + aggregation-AMG V-cycle (PCG_AGGREGATION_JACOBI.json)"
        else:
            w = f"7-pt Poisson {nx}x{nx}x{nx * world} fp64, PCG + aggregation-AMG V-cycle (PCG_AGGREGATION_JACOBI.json)"
    elif args.workload == "classical":
        w = f"7-pt Poisson {nx}^3 fp64, FGMRES(20) + classical AMG V-cycle (FGMRES_CLASSICAL_AGGRESSIVE_PMIS.json: PMIS, aggressive level 0, D2, JACOBI_L1 2 + 2)"
    elif args.workload == "banded":
        w = f"SuiteSparse-shaped banded-random CSR (nonsymmetric), {args.rows} rows, row length 3+Poisson(12), sigma 2000, fp64, FGMRES(20) + aggregation-AMG V-cycle"
    else:
        w = f"block 4x4 elasticity-like {nx}^3 block rows, {args.mode}, AMG V-cycle + MULTICOLOR_DILU (AGGREGATION_DILU)"
    return {"workload": w, "grid": nx}


# ------------------------------------------------------------------------------------------------------------------------
# reference GPU build beside it (context): oracle/_ref/ref_dump on the same generated matrix and configuration
# ------------------------------------------------------------------------------------------------------------------------
def write_ref_system(path, rp, ci, va, rhs, block=1):
    """a (block-)CSR system in the input format of oracle/ref_build/ref_dump.cu (header n, nnz, bx, by, has_diag, has_x0; then the arrays,
    values and vectors as doubles: the harness converts to the mode's precisions itself)"""
    with open(path, "wb") as f:
        f.write(np.array([rp.shape[0] - 1, ci.shape[0], block, block, 0, 0], np.int32).tobytes())
        for a, t in ((rp, np.int32), (ci, np.int32), (va, np.float64), (rhs, np.float64)):
            f.write(np.ascontiguousarray(a, t).tobytes())


def reference_gpu(nx, reps=2, config=None, timeout=300.0, system=None, mode="dDDI"):
    """config: a configuration dictionary (written to a file for the harness); default = PCG_AGGREGATION_JACOBI.json.
    system: a file written by write_ref_system instead of the generated nx^3 Poisson matrix."""
    exe = ROOT / "oracle" / "_ref" / "ref_dump"
    if not exe.exists():
        return {"unavailable": "oracle/_ref/ref_dump not built (oracle/ref_build/Makefile, needs /root/reference)"}
    try:
        cfg_path = str(CONFIG)
        if config is not None:
            cfg_path = "/tmp/amgxb_refgpu_cfg.json"
            Path(cfg_path).write_text(json.dumps(config))
        env = dict(os.environ, REFDUMP_NO_LEVELS="1", LD_LIBRARY_PATH=str(exe.parent) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([str(exe), system or f"poisson:{nx}", cfg_path, "/tmp/amgxb_refdump.bin", mode, str(reps)], capture_output=True, text=True, timeout=timeout, env=env)
        m = re.search(r"ref_dump: status (\d+) iterations (\d+) setup ([0-9.eE+-]+) s solve ([0-9.eE+-]+) s", r.stdout)
        if not m:
            return {"unavailable": "ref_dump gave no timing line", "tail": (r.stdout + r.stderr)[-300:]}
        st, it, ts, tsol = int(m.group(1)), int(m.group(2)), float(m.group(3)), float(m.group(4))
        return {"value": it / tsol, "unit": UNIT, "iterations": it, "solve_seconds": tsol, "setup_seconds": ts, "status": st,
                "how": f"unmodified reference (sm_100 build, oracle/_ref/libamgx_ref.so) through its own C API, best of {reps} AMGX_solver_solve calls timed with "
                       f"cudaEvents by oracle/ref_build/ref_dump.cu on the same GPU after this engine's timed region"}
    except Exception as e:      # never let the context figure cost the bench line
        return {"unavailable": repr(e)}


# ------------------------------------------------------------------------------------------------------------------------
# the reference's OWN CPU path beside the engine: the one solver family it implements completely on the host
# ------------------------------------------------------------------------------------------------------------------------
HOST_PATH_CFG = {"config_version": 2, "determinism_flag": 1, "solver": {
    "scope": "main", "solver": "PCG", "max_iters": 10, "monitor_residual": 1, "store_res_history": 1, "convergence": "RELATIVE_INI", "tolerance": 1e-30, "norm": "L2",
    "preconditioner": {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}}}


def reference_host_path(capi, rsc, nx, timeout=240.0):
    """BASELINE configs[0]'s solver (PCG + BLOCK_JACOBI) is the one whose every component the reference implements on the host
    (src/multiply.cu:753-852, src/solvers/block_jacobi_solver.cu:1256-1332; aggregation SIZE_2 and DILU have no host path).
    Here: that solver, 10 iterations, on the nx^3 Poisson matrix -- (a) the UNMODIFIED reference in its host mode hDDI through
    oracle/_ref/ref_dump (single-threaded C++ loops: cores = 1), (b) this engine on the same configuration and matrix.
    First written after round 2's GPU minutes were spent: the reference side reports `unavailable` with the reason if its host mode fails."""
    out = {"config": f"PCG + BLOCK_JACOBI(0.8), {HOST_PATH_CFG['solver']['max_iters']} iterations, 7-pt Poisson {nx}^3 fp64, b = 1, x0 = 0", "unit": UNIT}
    try:
        cfg = capi.Config(HOST_PATH_CFG)
        A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
        A.generate_poisson7(b, x, nx, nx, nx, 1, 1, 1)
        n, _, _ = A.get_size()
        slv = capi.Solver(rsc, cfg)
        slv.setup(A)
        tot_s, tot_it = 0.0, 0
        for i in range(6):                                  # 3 warm-up solves, 3 timed
            x.set_zero(n, 1)
            slv.solve(b, x, zero_initial_guess=True)
            if i >= 3:
                tot_s += slv.last_solve_stats()[0]
                tot_it += slv.iterations_number
        out["engine"] = {"value": tot_it / tot_s, "iterations": tot_it // 3, "how": "this engine, device-timed like `value`"}
        for o in (slv, x, b, A, cfg):
            o.destroy()
    except Exception as e:      # never let a context figure cost the bench line
        out["engine"] = {"unavailable": repr(e)}
    exe = ROOT / "oracle" / "_ref" / "ref_dump"
    if not exe.exists():
        out["reference_cpu"] = {"unavailable": "oracle/_ref/ref_dump not built"}
        return out
    try:
        cfg_path = "/tmp/amgxb_host_path_cfg.json"
        Path(cfg_path).write_text(json.dumps(HOST_PATH_CFG))
        env = dict(os.environ, REFDUMP_NO_LEVELS="1", LD_LIBRARY_PATH=str(exe.parent) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([str(exe), f"poisson:{nx}", cfg_path, "/tmp/amgxb_refdump_host.bin", "hDDI", "1"], capture_output=True, text=True, timeout=timeout, env=env)
        m = re.search(r"ref_dump: status (\d+) iterations (\d+) setup ([0-9.eE+-]+) s solve ([0-9.eE+-]+) s", r.stdout)
        if not m:
            out["reference_cpu"] = {"unavailable": "ref_dump (hDDI) gave no timing line", "tail": (r.stdout + r.stderr)[-300:]}
        else:
            it, ts, tsol = int(m.group(2)), float(m.group(3)), float(m.group(4))
            out["reference_cpu"] = {"value": it / tsol, "kind": "reference", "cores": 1, "iterations": it, "solve_seconds": tsol, "setup_seconds": ts,
                                    "how": "unmodified reference (oracle/_ref/libamgx_ref.so), mode hDDI = its own host implementation (single-threaded), one "
                                           "AMGX_solver_solve timed by oracle/ref_build/ref_dump.cu"}
    except Exception as e:
        out["reference_cpu"] = {"unavailable": repr(e)}
    return out


# ------------------------------------------------------------------------------------------------------------------------
# N > 1: parity of the distributed path against a single-rank solve of the assembled matrix (reduced size)
# ------------------------------------------------------------------------------------------------------------------------
def distributed_parity(capi, dist, torch, rsc, rank, world, local_rank):
    """Global grid g x g x (g*world), g = 48.  (a) distributed SpMV == single-rank SpMV of the assembled matrix, bit for bit;
    (b) PCG + BLOCK_JACOBI (no hierarchy: only reduction order differs): residual histories to 1e-12;
    (c) PCG + aggregation AMG: the hierarchy is built per slab (aggregates never cross partitions, as in the reference), so it is
        a different preconditioner than the single-rank one: iteration counts of both and the final true residual are reported."""
    from amgx_b200 import gallery
    g = 48
    nzg = g * world
    out = {"grid": [g, g, nzg]}
    jac = {"config_version": 2, "determinism_flag": 1, "solver": {"scope": "main", "solver": "PCG", "max_iters": 40, "monitor_residual": 1, "store_res_history": 1,
           "convergence": "RELATIVE_INI", "tolerance": 1e-10, "norm": "L2",
           "preconditioner": {"scope": "jac", "solver": "BLOCK_JACOBI", "relaxation_factor": 0.8, "max_iters": 1, "monitor_residual": 0}}}
    xg = np.random.default_rng(7).standard_normal(g * g * nzg)
    nloc = g * g * g
    res = {}
    for name, cfgobj in (("jacobi", capi.Config(jac)), ("amg", capi.Config(file=str(CONFIG)))):
        A, b, x, y = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc), capi.Vector(rsc)
        A.generate_poisson7(b, x, g, g, g, 1, 1, world)
        b.bind(A)
        x.bind(A)
        y.bind(A)
        if name == "jacobi":
            x.upload(xg[rank * nloc:(rank + 1) * nloc])
            y.set_zero(nloc)
            A.multiply(x, y)
            yl = y.download()
        x.set_zero(nloc)            # a zero initial guess is a PROMISE to the solver (as in the reference): the generator filled x with ones
        slv = capi.Solver(rsc, cfgobj)
        slv.setup(A)
        slv.solve(b, x, zero_initial_guess=True)
        hist = np.array(slv.residual_history())
        xl = x.download()
        parts = [torch.zeros(nloc, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(xl).cuda())
        xfull = torch.cat(parts).cpu().numpy()
        if name == "jacobi":
            parts = [torch.zeros(nloc, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(yl).cuda())
            yfull = torch.cat(parts).cpu().numpy()
        res[name] = (hist, slv.iterations_number, slv.status, xfull)
        for o in (slv, y, x, b, A, cfgobj):
            o.destroy()
    if rank == 0:
        cfg1 = capi.Config(jac)
        rsc1 = capi.Resources(cfg1, device=local_rank)
        rp, ci, va = gallery.poisson7pt(g, g, nzg)
        ng = rp.shape[0] - 1
        for name, cfgobj in (("jacobi", cfg1), ("amg", capi.Config(file=str(CONFIG)))):
            A = capi.Matrix(rsc1).upload(rp, ci, va)
            b = capi.Vector(rsc1).upload(np.ones(ng))
            x = capi.Vector(rsc1)
            if name == "jacobi":
                x.upload(xg)
                y = capi.Vector(rsc1).set_zero(ng)
                A.multiply(x, y)
                out["spmv_bit_exact"] = bool(np.array_equal(y.download(), yfull))
                y.destroy()
            x.set_zero(ng)
            slv = capi.Solver(rsc1, cfgobj)
            slv.setup(A)
            slv.solve(b, x, zero_initial_guess=True)
            h1 = np.array(slv.residual_history())
            hist, it, status, xfull = res[name]
            m = min(len(h1), len(hist))
            import scipy.sparse as sp
            true_res = float(np.linalg.norm(np.ones(ng) - sp.csr_matrix((va, ci, rp), shape=(ng, ng)) @ xfull))
            out[name] = {"iterations_distributed": int(it), "iterations_single_rank": int(slv.iterations_number), "status": status,
                         "max_rel_history_deviation": float(np.max(np.abs(h1[:m] - hist[:m]) / h1[0])),
                         "true_residual_of_distributed_solution_rel": true_res / float(h1[0]), "reported_final_residual_rel": float(hist[-1] / hist[0])}
            for o in (slv, x, b, A):
                o.destroy()
            if cfgobj is not cfg1:
                cfgobj.destroy()
        rsc1.destroy()
        cfg1.destroy()
        j = out.get("jacobi", {})
        out["green"] = bool(out.get("spmv_bit_exact") and j.get("max_rel_history_deviation", 1) < 1e-12 and
                            j.get("iterations_distributed") == j.get("iterations_single_rank") and out["amg"]["status"] == "success" and
                            abs(out["amg"]["true_residual_of_distributed_solution_rel"] - out["amg"]["reported_final_residual_rel"]) < 1e-9)
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------------------------------
# the other BASELINE workloads beside the default line (N = 1): child processes of this script, bounded in time
# ------------------------------------------------------------------------------------------------------------------------
_CURRENT_CHILD = [None]        # process group of the context child that is running (ended with the bench if the bench is terminated)


def _end_current_child():
    pg = _CURRENT_CHILD[0]
    if pg:
        try:
            os.killpg(pg, 9)
        except OSError:
            pass
    _CURRENT_CHILD[0] = None


EXTRA_WORKLOADS = [
    # most important first: a child that no longer fits the common time budget is skipped.  "+reference-gpu": the child also times the
    # unmodified reference GPU build on the same matrix and configuration (its reference_gpu object); the 512^3 aggregation problem does
    # not (30 s of reference solves; the main line carries that comparison at 256^3 on the same code path)
    ("poisson512", ["--workload", "poisson", "--grid", "512", "--steps", "2", "--warmup", "3"]),                                   # north star: >= 70 % of the roofline at 512^3
    ("banded4m", ["--workload", "banded", "--steps", "3", "--warmup", "3", "+reference-gpu"]),                                      # SURVEY 8(d) input 2
    ("classical512", ["--workload", "classical", "--grid", "512", "--steps", "2", "--warmup", "3", "+reference-gpu"]),              # BASELINE configs[2]
    ("block160_dDFI", ["--workload", "block", "--mode", "dDFI", "--steps", "3", "--warmup", "3", "+reference-gpu"]),                # BASELINE configs[4] at 160^3 block rows
]


def other_workloads(budget_s=240.0, per_run_s=110.0, workloads=None, script=None):
    """Runs `bench.py <flags> --no-cpu-baseline [--no-reference-gpu] --no-extras` once per extra workload and returns their JSON lines
    (None-valued keys dropped).  A child that fails, prints no line or runs out of time costs only its own entry."""
    out = {}
    t_end = time.time() + budget_s
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK")}
    for name, flags in (EXTRA_WORKLOADS if workloads is None else workloads):
        left = t_end - time.time()
        if left < 40.0:
            out[name] = {"skipped": "time budget of the extra workloads spent"}
            continue
        cmd = [sys.executable, str(script or (ROOT / "bench.py")), *[f for f in flags if f != "+reference-gpu"], "--no-cpu-baseline", "--no-extras"]
        if "+reference-gpu" not in flags:
            cmd.append("--no-reference-gpu")
        t0 = time.time()
        try:
            # own session: a child that overruns is ended together with whatever it started (the reference harness), so that nothing of it
            # is still on the GPU when the next child is timed
            proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=str(ROOT), start_new_session=True)
            _CURRENT_CHILD[0] = proc.pid
            try:
                c_out, c_err = proc.communicate(timeout=min(left, per_run_s))
            except subprocess.TimeoutExpired:
                _end_current_child()
                proc.communicate()
                raise
            finally:
                _CURRENT_CHILD[0] = None
            lines = [ln for ln in c_out.splitlines() if ln.startswith("{")]
            if not lines:
                out[name] = {"error": "no JSON line", "returncode": proc.returncode, "tail": (c_err or c_out)[-300:]}
                continue
            d = json.loads(lines[-1])
            d = {k: v for k, v in d.items() if v is not None}
            d["wall_seconds_of_the_child"] = time.time() - t0
            out[name] = d
        except subprocess.TimeoutExpired:
            out[name] = {"error": "timed out after %.0f s" % (time.time() - t0)}
        except Exception as e:      # never let an extra cost the bench line
            out[name] = {"error": repr(e)}
    return out


class LineGuard:
    """Holds rank 0's JSON line while the context objects of a multi-GPU run (parity, strong_512) are computed.  Those run collectives;
    should one of them hang, the guard prints the line as it stands -- its headline numbers are final before the guard is armed --
    and ends the process, on every rank, instead of leaving the launcher waiting."""

    def __init__(self, out, armed, limit_s):
        self.out = out
        self.lock = threading.Lock()
        self.done = False
        self.timer = None
        if armed:
            self.timer = threading.Timer(limit_s, self._expire)
            self.timer.daemon = True
            self.timer.start()
        self.limit_s = limit_s

    def _emit(self, note=None):
        with self.lock:
            if self.done:
                return False
            self.done = True
        if self.out is not None:
            if note:
                self.out["note"] = note
            print(json.dumps(self.out, default=str), flush=True)
        return True

    def _expire(self):
        if self._emit("context objects (parity / strong_512 / block_weak) did not finish within %.0f s: line printed without what was missing" % self.limit_s):
            sys.stdout.flush()
            os._exit(0)

    def finish(self):
        if self.timer is not None:
            self.timer.cancel()
        self._emit()


def strong_512(capi, dist, torch, rsc, cfg, world, steps, warmup, grid=512):
    """BASELINE configs[3] as stated: the grid^3 Poisson problem row-partitioned over the ranks (z-slabs of grid / world planes), PCG +
    aggregation AMG, zero initial guess; timed like the main line (the library's CUDA events on its solve stream, max over ranks)."""
    t0 = time.time()
    A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
    A.generate_poisson7(b, x, grid, grid, grid // world, 1, 1, world)
    n, _, _ = A.get_size()
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    b.bind(A)
    x.bind(A)
    t_setup = time.time() - t0

    def one():
        x.set_zero(n, 1)
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        return s, k, slv.iterations_number

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    tot_s, tot_k, tot_it = 0.0, 0, 0
    for _ in range(steps):
        s, k, it = one()
        tot_s += s
        tot_k += k
        tot_it += it
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([tot_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_s = float(t.item())
    status = slv.status
    hist = np.array(slv.residual_history()).ravel() if dist.get_rank() == 0 else np.zeros(0)      # read where the main line reads it
    res = {"workload": f"7-pt Poisson {grid}^3 fp64 row-partitioned over {world} GPUs (z-slabs of {grid // world} planes), PCG + aggregation-AMG V-cycle (PCG_AGGREGATION_JACOBI.json)",
           "value": tot_it / tot_s, "unit": UNIT, "scaling": "strong", "steps": steps, "warmup": warmup, "ms_per_step": tot_s / steps * 1e3,
           "iterations_per_step": tot_it / steps, "solve_status": status, "rows": n * world, "setup_seconds": t_setup, "gpu_launches": int(tot_k),
           "final_relative_residual": float(hist[-1] / hist[0]) if len(hist) else None,
           "n1_counterpart": "other_workloads.poisson512.value of the N=1 line (same grid, same solver, one GPU)"}
    for o in (slv, x, b, A):
        o.destroy()
    return res


def block_weak(capi, dist, torch, rsc, rank, world, nx=160, mode="dDFI", steps=3, warmup=3):
    """BASELINE configs[4]-style weak scaling: nx^3 4x4 block rows per GPU (z-slabs of the nx x nx x (nx * N) grid), AMG V-cycle +
    MULTICOLOR_DILU, fp32 matrix / fp64 vectors, uploaded through AMGX_matrix_upload_distributed (partition offsets) as an application
    would; timed like the main line.  The N = 1 counterpart is other_workloads.block160_dDFI of the N = 1 line."""
    import ctypes as C
    from amgx_b200 import gallery
    t0 = time.time()
    cfg = capi.Config(BLOCK_CFG)
    nz = nx * world
    lrp, lci, lva = gallery.block_elasticity_slab(nx, nx, nz, nx * rank, nx * (rank + 1), dtype=np.float32 if mode[2] == "F" else np.float64)
    n = lrp.shape[0] - 1
    lib = capi.load_library()
    A = capi.Matrix(rsc, mode)
    offsets = np.array([nx * nx * nx * r for r in range(world + 1)], np.int64)
    dh = C.c_void_p()
    if lib.AMGX_distribution_create(C.byref(dh), cfg.h) != 0 or lib.AMGX_distribution_set_partition_data(dh, 1, offsets.ctypes.data) != 0:
        raise RuntimeError("AMGX_distribution_create / set_partition_data failed")
    rc = lib.AMGX_matrix_upload_distributed(A.h, nx * nx * nz, n, lci.shape[0], 4, 4, lrp.ctypes.data, lci.ctypes.data, lva.ctypes.data, None, dh)
    lib.AMGX_distribution_destroy(dh)
    if rc != 0:
        raise RuntimeError(f"AMGX_matrix_upload_distributed returned {rc}")
    del lva, lci
    b, x = capi.Vector(rsc, mode), capi.Vector(rsc, mode)
    b.bind(A)
    x.bind(A)
    b.upload(np.ones(n * 4), block_dim=4)
    x.set_zero(n, 4)
    slv = capi.Solver(rsc, cfg, mode)
    slv.setup(A)
    t_setup = time.time() - t0

    def one():
        x.set_zero(n, 4)
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        return s, k, slv.iterations_number

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    tot_s, tot_k, tot_it = 0.0, 0, 0
    for _ in range(steps):
        s, k, it = one()
        tot_s += s
        tot_k += k
        tot_it += it
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([tot_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_s = float(t.item())
    res = {"workload": f"block 4x4 elasticity-like, {nx}^3 block rows per GPU x {world} GPUs, {mode}, AMG V-cycle + MULTICOLOR_DILU (AGGREGATION_DILU), AMGX_matrix_upload_distributed",
           "value": tot_it / tot_s * world, "unit": UNIT, "scaling": "weak", "global_iterations_per_sec": tot_it / tot_s, "steps": steps, "warmup": warmup,
           "ms_per_step": tot_s / steps * 1e3, "iterations_per_step": tot_it / steps, "solve_status": slv.status, "block_rows_global": nx * nx * nz,
           "generate_upload_setup_seconds": t_setup, "gpu_launches": int(tot_k),
           "n1_counterpart": "other_workloads.block160_dDFI.value of the N=1 line (same rows per GPU, same solver)"}
    for o in (slv, x, b, A, cfg):
        o.destroy()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson", choices=["poisson", "classical", "banded", "block"])
    ap.add_argument("--grid", dest="n", type=int, default=None, help="grid points per dimension (poisson: per rank unless --strong; block: block rows per dimension)")
    ap.add_argument("--rows", type=int, default=4_000_000, help="banded workload: number of rows")
    ap.add_argument("--mode", default="dDFI", choices=["dDDI", "dDFI", "dFFI"], help="block workload: AMGX mode")
    ap.add_argument("--strong", action="store_true", help="N > 1: keep the global grid at --grid^3 (BASELINE configs[3] with --grid 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the context runs: other_workloads and reference_host_path (N = 1), strong_512 and block_weak (N > 1)")
    args = ap.parse_args()
    if args.n is None:
        args.n = 160 if args.workload == "block" else 256
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload != "poisson":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm exists for the poisson workload (BASELINE configs[1])"}), flush=True)
            return
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from amgx_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed and args.workload != "poisson":
        raise SystemExit("multi-GPU bench lines exist for the poisson workload; see tools/bench_block_dist.py for the block configuration")
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the peer-memory receive window is bump-allocated per Resources and never shrinks: the main problem, the parity problems, strong_512
        # and block_weak all carve their levels out of it (about 30 MB together; default window 64 MB) -- leave room so that none of them
        # falls back to the NCCL path half way
        os.environ.setdefault("AMGXB_P2P_WINDOW_MB", "256")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    capi.initialize()
    capi.register_print_callback(None)
    mode = "dDDI"
    if args.workload == "block":
        mode = args.mode
        cfg = capi.Config(BLOCK_CFG)
    elif args.workload == "banded":
        cfg = capi.Config(BANDED_CFG)
    elif args.workload == "classical":
        cfg = capi.Config(CLASSICAL_CFG)
    else:
        cfg = capi.Config(file=str(CONFIG))
    comm = None
    if distributed:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.AMGXB200_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
    rsc = capi.Resources(cfg, device=local_rank, comm=comm) if (distributed or local_rank) else capi.Resources(cfg)
    A, b, x = capi.Matrix(rsc, mode), capi.Vector(rsc, mode), capi.Vector(rsc, mode)
    nx = args.n
    bd = 1
    t0 = time.time()
    if args.workload == "classical":
        A.generate_poisson7(b, x, nx, nx, nx, 1, 1, 1)
    elif args.workload == "poisson":
        if args.strong:
            if nx % world:
                raise SystemExit("--strong needs --grid divisible by the number of ranks")
            A.generate_poisson7(b, x, nx, nx, nx // world, 1, 1, world)      # z-slabs of the fixed nx^3 grid
        else:
            A.generate_poisson7(b, x, nx, nx, nx, 1, 1, world)               # z-slabs: global grid nx * nx * (nx*world)
    elif args.workload == "banded":
        from amgx_b200 import gallery
        rp, ci, va = gallery.random_banded(args.rows)
        A.upload(rp, ci, va)
        b.upload(np.ones(rp.shape[0] - 1))
        if rank == 0 and not args.no_reference_gpu:      # the same matrix for the reference GPU build, run after this engine's timed region
            try:
                write_ref_system(REF_SYSTEM, rp, ci, va, np.ones(rp.shape[0] - 1))
            except Exception:
                pass
        del rp, ci, va
    else:
        from amgx_b200 import gallery
        bd = 4
        rp, ci, va = gallery.block_elasticity(nx, nx, nx, dtype=np.float32 if mode[2] == "F" else np.float64)
        A.upload(rp, ci, va, block_dims=(4, 4))
        b.upload(np.ones((rp.shape[0] - 1) * 4), block_dim=4)
        if rank == 0 and not args.no_reference_gpu:      # the same matrix for the reference GPU build (3.7 GB of doubles at 160^3), run after this engine's timed region
            try:
                write_ref_system(REF_SYSTEM, rp, ci, va, np.ones((rp.shape[0] - 1) * 4), block=4)
            except Exception:
                try:
                    os.remove(REF_SYSTEM)
                except OSError:
                    pass
        del rp, ci, va
    n, _, _ = A.get_size()
    nnz = A.get_nnz()
    slv = capi.Solver(rsc, cfg, mode)
    slv.setup(A)
    t_setup = time.time() - t0
    if distributed:
        b.bind(A)
        x.bind(A)

    def sync_all():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ----
    def one_solve():
        x.set_zero(n, bd)
        slv.solve(b, x, zero_initial_guess=True)
        s, k = slv.last_solve_stats()
        return s, k, slv.iterations_number

    for _ in range(args.warmup):
        one_solve()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    tot_s, tot_k, tot_it = 0.0, 0, 0
    for _ in range(args.steps):
        s, k, it = one_solve()
        tot_s += s
        tot_k += k
        tot_it += it
    sync_all()
    clocks = sampler.stop()
    if distributed:
        t = torch.tensor([tot_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_s = float(t.item())
    # weak scaling: every rank owns one N=1 workload, so the job advances `world` N=1-sized problems by one V-cycle iteration per outer
    # iteration: value = iterations/s x world; the raw rate of the (world x larger) global problem is config.global_iterations_per_sec.
    # strong scaling: the global problem is the N=1 problem: value = raw rate.
    raw_its = tot_it / tot_s
    value = raw_its if (args.strong or not distributed) else raw_its * world
    status = slv.status
    hist = np.array(slv.residual_history()).ravel() if rank == 0 else None

    # ---- end to end through the C-ABI with host buffers ----
    dt_v = np.float64 if mode[1] == "D" else np.float32
    hb = np.ones(n * bd, dtype=dt_v)
    hx = np.zeros(n * bd, dtype=dt_v)
    lib = capi.load_library()
    lib.AMGX_pin_memory(hb.ctypes.data, hb.nbytes)
    lib.AMGX_pin_memory(hx.ctypes.data, hx.nbytes)

    def one_e2e():
        b.upload(hb, block_dim=bd)
        x.set_zero(n, bd)
        slv.solve(b, x, zero_initial_guess=True)
        x.download(hx)
        return slv.iterations_number

    one_e2e()
    sync_all()
    t0 = time.perf_counter()
    e_it = 0
    for _ in range(args.steps):
        e_it += one_e2e()
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
    lib.AMGX_unpin_memory(hb.ctypes.data)
    lib.AMGX_unpin_memory(hx.ctypes.data)
    e2e_value = e_it / e_dt * (1 if (args.strong or not distributed) else world)

    # ---- roofline of the dominant kernel ----
    roof = None
    if not distributed:
        try:
            peak, peak_src = measured_peak()
            msz = 8 if mode[2] == "D" else 4
            vsz = 8 if mode[1] == "D" else 4
            ms = A.bench_kernel(0, warmup=3, reps=20, flush_l2=False)      # operands (>= 1.4 GB) >> 126 MB L2
            byt = nnz * (msz * bd * bd + 4) + n * 4
            enc = os.environ.get("AMGXB_COLENC", "")
            kinfo = A.kernel_info() if bd == 1 else {}
            if kinfo.get("window"):
                family = "csr_window_kernel<%s> (x window of %d entries in shared memory, 16-bit column offsets)"
                family = family.replace("%d", str(kinfo["window"]))
            elif kinfo.get("coded_tiles"):
                family = "csr_tile_enc_kernel<%s> (coded column / value streams)"
            else:
                family = "csr_tile_kernel<%s>"
            spmv = {"kernel": (family % "EPI_SPMV" if bd == 1 else "block4_tile_kernel<SPMV> (TMA-staged 4x4 blocks)") + " (fine level)", "ms_per_launch": ms, "algorithmic_bytes": byt,
                    "achieved": byt / ms / 1e6, "frac": byt / ms / 1e6 / peak}
            if bd == 1:
                ms_j = A.bench_kernel(1, warmup=3, reps=20, flush_l2=False)
                byt_j = byt + 4 * n * vsz
                traffic = None
                tf = ROOT / "profiles" / (f"r02_ncu_traffic_jacobi_{nx}.json" if args.workload == "poisson" else "r02_ncu_traffic_jacobi_banded.json")
                if tf.exists() and (args.workload == "poisson" or (args.workload == "banded" and kinfo.get("window"))):
                    traffic = json.loads(tf.read_text()).get("traffic_bytes_per_launch")
                roof = {"bound": "hbm", "achieved": byt_j / ms_j / 1e6, "peak": peak, "unit": "GB/s", "frac": byt_j / ms_j / 1e6 / peak, "traffic": traffic,
                        "kernel": "fused Jacobi sweep, fine level: " + family % "EPI_JACOBI", "kernel_plan": kinfo,
                        "ms_per_launch": ms_j, "algorithmic_bytes": byt_j, "peak_source": peak_src,
                        "share_of_iteration": "~70 % (profiles/r02_launches_solve_256.md)" if args.workload == "poisson" else None,
                        "spmv": spmv}
                if traffic:     # the same launch by the bytes that actually crossed the HBM interface (ncu dram__bytes_read + write, committed capture)
                    roof["dram"] = {"bytes_per_launch": traffic, "achieved": traffic / ms_j / 1e6, "frac": traffic / ms_j / 1e6 / peak,
                                    "note": "coded matrix streams move fewer bytes than the north-star formula charges, so `frac` above can exceed 1; this is the DRAM-side view of the same time"}
            else:
                roof = {"bound": "hbm", "achieved": spmv["achieved"], "peak": peak, "unit": "GB/s", "frac": spmv["frac"], "traffic": None, "kernel": spmv["kernel"],
                        "ms_per_launch": ms, "algorithmic_bytes": byt, "peak_source": peak_src}
            # whole outer iteration against the same peak: SURVEY 8(d)'s per-unit figures summed over the hierarchy the setup actually built
            # (PCG outside M^-1: M(A_0) + 12 N 8; per level: 3 fused post-sweeps M(A_l) + 4 n_l 8 each, restriction and prolongation
            # n_l (4 + 8) + n_{l+1} 8 each; presweeps = 0 and a zero initial guess leave no residual pass; coarsest: a zero-guess sweep + a full one)
            if args.workload == "poisson":
                try:
                    lv = [slv.level_info(l) for l in range(slv.num_levels())]
                    M = lambda i: i["nnz"] * 12 + i["n"] * 4
                    it_bytes = M(lv[0]) + 12 * lv[0]["n"] * 8
                    for l, i in enumerate(lv):
                        if l + 1 < len(lv):
                            it_bytes += 3 * (M(i) + 4 * i["n"] * 8) + 2 * (i["n"] * 12 + lv[l + 1]["n"] * 8)
                        else:
                            it_bytes += 3 * i["n"] * 8 + (M(i) + 4 * i["n"] * 8)
                    ms_it = tot_s / max(tot_it, 1) * 1e3
                    roof["iteration"] = {"algorithmic_bytes": int(it_bytes), "levels": len(lv), "ms_at_peak": it_bytes / peak / 1e6, "ms_measured": ms_it,
                                         "achieved": it_bytes / ms_it / 1e6, "frac": it_bytes / ms_it / 1e6 / peak,
                                         "operator_complexity": sum(i["nnz"] for i in lv) / lv[0]["nnz"]}
                except Exception as e:      # never let the extra figure cost the bench line
                    roof["iteration"] = {"error": repr(e)}
        except Exception as e:      # a kernel-level figure must not cost the line its solve-level numbers
            roof = {"error": repr(e)}

    # ---- the line: value, e2e, roofline and clocks are final here; what follows only adds context objects to it ----
    out = None
    if rank == 0:
        cfgd = workload_config(args, world)
        cfgd.update({"rows": n * world, "nnz_per_rank": nnz, "iterations_per_step": tot_it / args.steps, "solve_status": status,
                     "global_iterations_per_sec": raw_its,
                     "value_definition": ("outer iterations (one V-cycle each) per second of the fixed global problem" if (args.strong or not distributed) else
                                          "outer PCG iterations (one V-cycle each) per second x number of N=1-sized sub-problems (= n_gpus)"),
                     "l2": "inputs larger than L2 (matrix alone %.2f GB per rank)" % (nnz * (12 if bd == 1 else 16 * (8 if mode[2] == 'D' else 4) + 4) / 1e9),
                     "setup_seconds": t_setup, "parallelism": f"row-partition x{world}" if distributed else "single GPU",
                     "exchange": ("peer-memory stores over NVLink (CUDA IPC)" if os.environ.get("AMGXB_P2P", "1") != "0" else "NCCL send/recv") if distributed else None})
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
               "dtype": "f64" if mode == "dDDI" else ("f32 matrix / f64 vectors" if mode == "dDFI" else "f32"), "data": "synthetic",
               "config": cfgd,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hb.nbytes), "d2h_bytes_per_step": int(hx.nbytes),
                       "ms_per_step": e_dt / args.steps * 1e3},
               "gpu_launches": int(tot_k), "clocks": clocks, "roofline": roof, "cpu_baseline": None, "reference_gpu": None, "parity": None,
               "final_relative_residual": float(hist[-1] / hist[0]) if hist is not None and len(hist) else None}
    guard = LineGuard(out, armed=distributed, limit_s=float(os.environ.get("AMGXB_BENCH_GUARD_S", "300")))      # parity and strong_512 run collectives: a hang there must not cost the line
    # the context objects of the N = 1 line are child processes with their own time limits; together they get AMGXB_BENCH_CONTEXT_S seconds,
    # and a SIGTERM from whoever launched the bench prints the line as it stands instead of losing it
    ctx_deadline = time.time() + float(os.environ.get("AMGXB_BENCH_CONTEXT_S", "360"))
    ctx_left = lambda cap: max(5.0, min(cap, ctx_deadline - time.time()))
    try:
        import signal
        signal.signal(signal.SIGTERM, lambda *_a: (_end_current_child(), guard._emit("terminated from outside while the context objects were running: line printed as it stood"), os._exit(0)))
    except Exception:
        pass

    if distributed and not args.no_parity:
        try:
            parity = distributed_parity(capi, dist, torch, rsc, rank, world, local_rank)
        except Exception as e:
            parity = {"error": repr(e)}
        if out is not None:
            out["parity"] = parity

    for o in (slv, x, b, A):
        o.destroy()

    if distributed and not args.no_extras and args.workload == "poisson" and not args.strong and nx == 256 and 512 % world == 0:
        try:
            torch.cuda.empty_cache()
            st = strong_512(capi, dist, torch, rsc, cfg, world, steps=3, warmup=3)
        except Exception as e:
            st = {"error": repr(e)}
        if out is not None:
            out["strong_512"] = st
        try:
            bw = block_weak(capi, dist, torch, rsc, rank, world, nx=int(os.environ.get("AMGXB_BENCH_BLOCK_NX", "160")))
        except Exception as e:
            bw = {"error": repr(e)}
        if out is not None:
            out["block_weak"] = bw

    if rank == 0 and not distributed and args.workload == "classical" and not args.no_reference_gpu:
        torch.cuda.empty_cache()
        out["reference_gpu"] = reference_gpu(nx, config=CLASSICAL_CFG, timeout=ctx_left(300.0))
    if rank == 0 and not distributed and args.workload in ("banded", "block") and not args.no_reference_gpu:
        torch.cuda.empty_cache()
        out["reference_gpu"] = reference_gpu(0, config=BANDED_CFG if args.workload == "banded" else BLOCK_CFG, timeout=ctx_left(300.0), system=REF_SYSTEM, mode=mode) \
            if os.path.exists(REF_SYSTEM) else {"unavailable": "the system file for the reference could not be written"}
        try:
            os.remove(REF_SYSTEM)
        except OSError:
            pass
    if rank == 0 and not distributed and args.workload == "poisson":
        if not args.no_reference_gpu:
            torch.cuda.empty_cache()
            out["reference_gpu"] = reference_gpu(nx, timeout=ctx_left(300.0))
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = oracle_baseline(nx, nx, nx, 5 if nx <= 256 else 2)
        if not args.no_extras and not args.no_reference_gpu and not args.no_cpu_baseline and nx == 256:
            out["reference_host_path"] = reference_host_path(capi, rsc, nx, timeout=ctx_left(240.0))
        if not args.no_extras and nx == 256:
            out["other_workloads"] = other_workloads(budget_s=ctx_left(240.0))
    guard.finish()
    # the line is out; a teardown that hangs (a communicator left in a bad state by a failed context object) must not keep the launcher waiting
    t_exit = threading.Timer(90.0, lambda: os._exit(0))
    t_exit.daemon = True
    t_exit.start()
    for o in (rsc, cfg):
        o.destroy()
    capi.finalize()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
