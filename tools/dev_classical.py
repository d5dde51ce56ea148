"""Development check (GPU): classical AMG of the engine vs the CPU oracle, level by level."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi, gallery  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.golden.make_golden import cfg_fgmres_classical  # noqa: E402

orc.set_num_threads(8)
capi.initialize()


def run(name, A3, **kw):
    rp, ci, va = A3
    n = rp.shape[0] - 1
    cfgd = cfg_fgmres_classical(**kw)
    amgc = cfgd["solver"]["preconditioner"]
    cfg = capi.Config(cfgd)
    rsc = capi.Resources(cfg)
    A = capi.Matrix(rsc).upload(rp, ci, va)
    b = capi.Vector(rsc).upload(np.ones(n))
    x = capi.Vector(rsc).set_zero(n)
    slv = capi.Solver(rsc, cfg)
    t = time.time()
    slv.setup(A)
    ts = time.time() - t
    slv.solve(b, x, zero_initial_guess=True)
    hist = slv.residual_history()
    nl = slv.num_levels()
    t = time.time()
    o = orc.ClassicalAMG(rp, ci, va, max_levels=50, min_coarse_rows=2, presweeps=amgc["presweeps"], postsweeps=amgc["postsweeps"], coarsest_sweeps=2,
                         smoother="JACOBI_L1", omega=1.0, strength_threshold=amgc["strength_threshold"], max_row_sum=amgc["max_row_sum"],
                         interpolator=amgc["interpolator"], aggressive_levels=amgc["aggressive_levels"], interp_max_elements=amgc["interp_max_elements"])
    to = time.time() - t
    ok = (nl == o.num_levels())
    msgs = []
    for l in range(min(nl, o.num_levels())):
        g = slv.level_matrix(l)
        ol = o.level(l)
        same = g[0].shape == ol["row_ptr"].shape and np.array_equal(g[0], ol["row_ptr"]) and np.array_equal(g[1], ol["col_idx"])
        vsame = same and np.array_equal(g[2], ol["values"])
        vdiff = float(np.max(np.abs(g[2] - ol["values"]) / np.maximum(np.abs(ol["values"]), 1e-300))) if same and g[2].size else -1
        m = f"L{l}: n={g[0].shape[0]-1}/{ol['n']} nnz={g[1].shape[0]}/{ol['nnz']} pattern={'=' if same else 'X'} values={'=' if vsame else ('%.1e' % vdiff)}"
        if "cf_map" in ol and l < nl - 1:
            cf = slv.level_cf_map(l)
            P = slv.level_P(l)
            cfs = np.array_equal(cf, ol["cf_map"])
            ps = np.array_equal(P[0], ol["P_row_offsets"]) and np.array_equal(P[1], ol["P_col_indices"])
            pv = ps and np.array_equal(P[2], ol["P_values"])
            m += f" cf={'=' if cfs else 'X'} P.pattern={'=' if ps else 'X'} P.values={'=' if pv else 'X'}"
            ok = ok and cfs and ps and pv
        ok = ok and same and vsame
        msgs.append(m)
    xo, ito, histo, convo = orc.fgmres(rp, ci, va, np.ones(n), amg=o, tol=cfgd["solver"]["tolerance"], max_iters=cfgd["solver"]["max_iters"],
                                        restart=cfgd["solver"]["gmres_n_restart"])
    k = min(len(hist), len(histo))
    hd = float(np.max(np.abs(hist[:k] - histo[:k]) / histo[0]))
    print(f"[{name}] levels {nl}/{o.num_levels()} iters {slv.iterations_number}/{ito} hist_dev {hd:.2e} final {hist[-1]/hist[0]:.2e} setup {ts:.2f}s (oracle {to:.2f}s) "
          f"{'OK' if ok and slv.iterations_number == ito else 'MISMATCH'}", flush=True)
    for m in msgs:
        print("    " + m)
    for obj in (slv, x, b, A, rsc, cfg):
        obj.destroy()


run("poisson12 aggr+trunc4", gallery.poisson7pt(12))
run("poisson16x12x9 aggr+trunc4", gallery.poisson7pt(16, 12, 9))
run("poisson12s d2 notrunc", gallery.poisson7pt_sorted(12), aggressive_levels=0, max_elements=-1)
run("banded3000 d2 trunc4", gallery.random_banded(3000, sigma=40.0), aggressive_levels=0, max_iters=40)
run("poisson40 aggr+trunc4", gallery.poisson7pt(40), tol=1e-8)
run("poisson64 multipass-all", gallery.poisson7pt(64), tol=1e-8, interpolator="MULTIPASS", aggressive_levels=2)
