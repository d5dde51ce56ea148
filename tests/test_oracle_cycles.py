"""CPU: W and F cycles of the oracle (src/cycles/{w,f}_cycle.cu over fixed_cycle.cu) -- composition rules and convergence."""
import numpy as np
import pytest

from amgx_b200 import gallery


def test_w_and_f_cycles_converge_faster_than_v(oracle):
    rp, ci, va = gallery.poisson7pt(18)
    n = rp.shape[0] - 1
    its = {}
    for cyc in "VWF":
        amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8).set_cycle(cyc)
        x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=amg, tol=1e-8, max_iters=100)
        assert conv
        assert np.linalg.norm(np.ones(n) - gallery.to_scipy(rp, ci, va) @ x) <= 1.01e-8 * np.sqrt(n)
        its[cyc] = it
    assert its["W"] <= its["F"] <= its["V"] and its["W"] < its["V"]


def test_two_level_hierarchy_all_cycles_coincide(oracle):
    """with a single coarse level the next level is always the coarsest: one fixed cycle whatever the type (fixed_cycle.cu:169-179)"""
    rp, ci, va = gallery.poisson7pt(8)
    n = rp.shape[0] - 1
    b = np.random.default_rng(1).standard_normal(n)
    out = []
    for cyc in "VWF":
        amg = oracle.AMG(rp, ci, va, max_levels=2, presweeps=1, postsweeps=2, omega=0.8).set_cycle(cyc)
        assert amg.num_levels() == 2
        out.append(amg.vcycle(b))
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


def test_w_cycle_on_three_levels_is_two_coarse_visits(oracle):
    """three levels: W visits level 1 twice (each visit one fixed cycle on level 2); F = W visit then V visit: identical here"""
    rp, ci, va = gallery.poisson7pt(10)
    n = rp.shape[0] - 1
    b = np.random.default_rng(2).standard_normal(n)
    res = {}
    for cyc in "VWF":
        amg = oracle.AMG(rp, ci, va, max_levels=3, presweeps=1, postsweeps=1, omega=0.8).set_cycle(cyc)
        assert amg.num_levels() == 3
        x = amg.vcycle(b)
        res[cyc] = np.linalg.norm(b - gallery.to_scipy(rp, ci, va) @ x)
    assert np.isclose(res["W"], res["F"], rtol=1e-14)      # on 3 levels the W and the V visit of level 1 are the same fixed cycle
    assert res["W"] < res["V"]
