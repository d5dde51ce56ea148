"""The components added after round 1's GPU minutes were spent, one case list (tests/golden/cases_round2.py) three ways:
  * CPU, always:   the oracle runs every case (convergence / sanity -- the algebra itself is pinned in test_oracle_*.py);
  * CPU, when tests/golden/r2_<case>.npz exists (reference run by make_golden.py): oracle == reference, iteration count and
    residual history to 1e-12 -- until then these restatements are "parity unpinned";
  * GPU (opt-in until validated): engine == oracle on every case, and engine == reference where the golden exists."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests._gpu_util import UNVALIDATED, run_engine
from tests.golden.cases_round2 import case_dict
from tests.oracle_from_config import run_oracle

GOLD = Path(__file__).parent / "golden"
CASES = case_dict()
# cases that do not (and must not) converge: a single GMRES iteration, the stale-x Chebyshev on the coarsest level
NOT_CONVERGING = {"poisson9_gmres_one_iteration", "poisson10_amg_agg_cheb2_coarsest1"}
# PARALLEL_GREEDY: the reference colours in place and racily, engine and oracle synchronously -- same rule, possibly another valid
# colouring, hence another (equally good) smoother: the golden is compared on convergence and iteration count within 2
REFERENCE_NONDETERMINISTIC = {"poisson16_fgmres_agg_dilu_pgreedy", "symbanded3000_fgmres_agg_dilu_pgreedy", "poisson14_amg_gs_pgreedy"}
# host-side dots of nearly cancelling quantities inside the cycle: one digit of slack
TOL = {"poisson15x12x10_pcgf_agg_CG": 1e-11, "poisson15x12x10_pcgf_agg_CGF": 1e-11, "poisson12_fgmres_agg_CG3": 1e-11, "poisson12_amg_classical_CG": 1e-11,
       "poisson14x12x9_amg_agg_size4_cgcycle": 1e-11,
       # restarted GMRES(7) on b = 1 (a highly symmetric right-hand side) runs 108 iterations through near-degenerate Krylov spaces: a 1-ulp
       # perturbation of b moves the ORACLE's own history by 1.1e-10 (measured), so no summation order other than cuBLAS's own can do better
       "poisson12_gmres_noprec": 1e-9, "poisson12_gmres_jacobi": 1e-9,
       # D1 + interp_max_elements on levels >= 1: four rows of P_1 hold two coarse points with EXACTLY equal weights at the truncation cut;
       # the reference keeps the one its SpGEMM hash table emitted first (A_1 has hash-ordered columns there, ours ascending ones:
       # DESIGN 3.5).  Same level sizes, same iteration count; the history differs in the 5th digit.  Parity for this case: partial.
       "poisson16x12x9_fgmres_classical_d1_trunc4": 2e-4}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_runs_case(oracle, name):
    (rp, ci, va), cfg = CASES[name]
    n = rp.shape[0] - 1
    x, it, hist, conv, amg = run_oracle(oracle, cfg, rp, ci, va, np.ones(n))
    assert conv == (name not in NOT_CONVERGING)
    assert np.all(np.isfinite(hist)) and len(hist) == it + 1


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(oracle, name):
    f = GOLD / f"r2_{name}.npz"
    if not f.exists():
        pytest.skip("no reference golden yet (gpurun -- 'python tests/golden/make_golden.py r2'): parity unpinned")
    d = np.load(f)
    cfg = json.loads(str(d["config_json"]))
    x, it, hist, conv, amg = run_oracle(oracle, cfg, d["sys_row_ptr"], d["sys_col_idx"], d["sys_values"], d["sys_rhs"])
    href = d["res_history"]
    if name in REFERENCE_NONDETERMINISTIC:
        assert conv == (int(d["status"][0]) == 0) and abs(it - int(d["iterations"][0])) <= 2
        return
    assert it == int(d["iterations"][0]) and conv == (int(d["status"][0]) == 0)
    assert np.max(np.abs(hist - href) / href[0]) < TOL.get(name, 1e-12)
    if amg is not None and "num_levels" in d:      # ref_dump finds the hierarchy only under PCG / FGMRES / AMG-as-main-solver
        assert amg.num_levels() == int(d["num_levels"][0])


@pytest.mark.gpu
@UNVALIDATED
@pytest.mark.parametrize("name", list(CASES))
def test_engine_matches_oracle_and_golden(amgx, oracle, name):
    (rp, ci, va), cfg = CASES[name]
    n = rp.shape[0] - 1
    b = np.ones(n)
    x, it, status, hist = run_engine(amgx, cfg, rp, ci, va, b)
    xo, ito, histo, convo, amg = run_oracle(oracle, cfg, rp, ci, va, b)
    assert it == ito and (status == "success") == convo
    assert np.max(np.abs(hist - histo) / histo[0]) < TOL.get(name, 1e-12)
    f = GOLD / f"r2_{name}.npz"
    if f.exists():
        d = np.load(f)
        if name in REFERENCE_NONDETERMINISTIC:
            assert abs(it - int(d["iterations"][0])) <= 2
            return
        assert it == int(d["iterations"][0])
        assert np.max(np.abs(hist - d["res_history"]) / d["res_history"][0]) < TOL.get(name, 1e-12)
