"""Is the aggregation setup bitwise reproducible run to run at scale?  (two independent setups of the same matrix)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amgx_b200 import capi
capi.initialize(); capi.register_print_callback(None)
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 160
cfg = capi.Config(file=str(ROOT / "amgx_b200" / "configs" / "PCG_AGGREGATION_JACOBI.json"))
rsc = capi.Resources(cfg)
A, b, x = capi.Matrix(rsc), capi.Vector(rsc), capi.Vector(rsc)
A.generate_poisson7(b, x, nx, nx, nx)
runs = []
for rep in range(3):
    slv = capi.Solver(rsc, cfg)
    slv.setup(A)
    nl = slv.num_levels()
    data = []
    for l in range(nl):
        info = slv.level_info(l)
        rec = {"n": info["n"], "nnz": info["nnz"]}
        if l < nl - 1:
            rec["agg"] = slv.level_aggregates(l)[0]
        if l > 0 and info["n"] < 3_000_000:
            rp, ci, va = slv.level_matrix(l)
            rec["rp"], rec["ci"], rec["va"] = rp, ci, va
        data.append(rec)
    runs.append(data)
    slv.destroy()
ok = True
for rep in (1, 2):
    for l, (a, c) in enumerate(zip(runs[0], runs[rep])):
        same_sz = (a["n"], a["nnz"]) == (c["n"], c["nnz"])
        msg = f"run0 vs run{rep} level {l}: sizes {'=' if same_sz else 'DIFFER'} ({a['n']},{a['nnz']}) vs ({c['n']},{c['nnz']})"
        if "agg" in a and same_sz:
            d = int(np.count_nonzero(a["agg"] != c["agg"]))
            msg += f" aggregates differing: {d}"
            ok &= d == 0
        if "va" in a and same_sz:
            e = np.array_equal(a["ci"], c["ci"]) and np.array_equal(a["va"], c["va"])
            msg += f" matrix bitwise equal: {e}"
            ok &= e
        ok &= same_sz
        if not same_sz or "DIFFER" in msg or "False" in msg or (("differing: 0" not in msg) and "agg" in a):
            print(msg)
print("DETERMINISTIC" if ok else "NOT DETERMINISTIC", "levels", len(runs[0]), [r["n"] for r in runs[0]][:8])
