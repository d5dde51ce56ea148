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


# ---------------------------------------------------------------------------------------------------------------
# CG / CGF cycles and error scaling: the C restatement against an independent numpy multilevel implementation
# built from the oracle's own level arrays (matrices, aggregates)
# ---------------------------------------------------------------------------------------------------------------
class NumpyAMG:
    """fixed_cycle.cu + cg_cycle.cu / cg_flex_cycle.cu + aggregation error scaling, dense numpy, Jacobi smoother"""

    def __init__(self, amg, pre, post, omega, coarsest_sweeps=2, cycle="V", cycle_iters=2, error_scaling=0, steps=2):
        self.lv = [amg.level(l) for l in range(amg.num_levels())]
        for L in self.lv:
            L["A"] = gallery.to_scipy(L["row_ptr"], L["col_idx"], L["values"])
        self.pre, self.post, self.omega, self.cs = pre, post, omega, coarsest_sweeps
        self.cycle, self.iters, self.es, self.steps = cycle, cycle_iters, error_scaling, steps

    def smooth(self, L, b, x, zero, sweeps):
        for it in range(sweeps):
            if it == 0 and zero:
                x = self.omega * b / L["d"]
            else:
                x = x + self.omega * (b - L["A"] @ x) / L["d"]
        return x

    def fixed(self, l, b, x, zero, typ):
        L = self.lv[l]
        last = l == len(self.lv) - 1
        npre = self.cs if last else self.pre
        if npre > 0:
            x = self.smooth(L, b, x, zero, npre)
        elif zero:
            x = np.zeros_like(b)
        if last:
            return x
        r = b - L["A"] @ x
        agg = L["aggregates"]
        bc = np.zeros(L["n_coarse"])
        np.add.at(bc, agg, r)
        nxt_last = l + 1 == len(self.lv) - 1
        if typ == "V" or nxt_last:
            xc = self.fixed(l + 1, bc, None, True, "V")
        elif typ == "W":
            xc = self.fixed(l + 1, bc, None, True, "W")
            xc = self.fixed(l + 1, bc, xc, False, "W")
        elif typ == "F":
            xc = self.fixed(l + 1, bc, None, True, "W")
            xc = self.fixed(l + 1, bc, xc, False, "V")
        else:
            xc = self.cg(l + 1, bc, typ == "CGF")
        e = xc[agg]
        if self.es >= 2:
            if self.steps > 0:
                e = self.smooth(L, r, e, False, self.steps)
            Ae = L["A"] @ e
            nom, den = (r @ Ae, Ae @ Ae) if self.es == 2 else (r @ e, e @ Ae)
            lam = nom / den if den != 0 else 1.0
            lam = np.sign(lam) * min(max(abs(lam), 0.3), 10.0)
            x = x + lam * e
        else:
            x = x + e
        if self.post > 0:
            x = self.smooth(L, b, x, False, self.post)
        return x

    def cg(self, l, b, flex):
        A = self.lv[l]["A"]
        typ = "CGF" if flex else "CG"
        x = np.zeros_like(b)
        r = b - A @ x
        z = self.fixed(l, r, None, True, typ)
        p = z.copy()
        rz = r @ z
        k = 0
        while True:
            y = A @ p
            if flex:
                rz = r @ z
            alpha = rz / (y @ p)
            x = x + alpha * p
            k += 1
            if k == self.iters:
                return x
            d = r.copy()
            r = r - alpha * y
            d = r - d
            z = self.fixed(l, r, None, True, typ)
            if flex:
                beta = (z @ d) / rz
            else:
                rz_old, rz = rz, r @ z
                beta = rz / rz_old
            p = z + beta * p

    def vcycle(self, b):
        return self.fixed(0, b, None, True, self.cycle)


@pytest.mark.parametrize("cyc,iters", [("V", 2), ("W", 2), ("F", 2), ("CG", 1), ("CG", 2), ("CG", 3), ("CGF", 2), ("CGF", 3)])
def test_cycles_match_numpy_multilevel(oracle, cyc, iters):
    rp, ci, va = gallery.poisson7pt(11, 9, 8)
    n = rp.shape[0] - 1
    b = np.random.default_rng(5).standard_normal(n)
    amg = oracle.AMG(rp, ci, va, max_levels=5, presweeps=1, postsweeps=2, omega=0.8).set_cycle(cyc).set_cycle_iters(iters)
    assert amg.num_levels() == 5
    x = amg.vcycle(b)
    ref = NumpyAMG(amg, 1, 2, 0.8, cycle=cyc, cycle_iters=iters).vcycle(b)
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("es", [2, 3])
@pytest.mark.parametrize("steps", [0, 2])
def test_error_scaling_matches_numpy_multilevel(oracle, es, steps):
    rp, ci, va = gallery.poisson7pt(12, 10, 7)
    n = rp.shape[0] - 1
    b = np.random.default_rng(6).standard_normal(n)
    amg = oracle.AMG(rp, ci, va, max_levels=6, presweeps=1, postsweeps=1, omega=0.8).set_error_scaling(es, steps, 0)
    x = amg.vcycle(b)
    ref = NumpyAMG(amg, 1, 1, 0.8, error_scaling=es, steps=steps).vcycle(b)
    assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_error_scaling_reuse_scale_skips_recomputation(oracle):
    """reuse_scale = k: the next k corrections of a level use the stored scale on the raw (unsmoothed) prolongation"""
    rp, ci, va = gallery.poisson7pt(10)
    n = rp.shape[0] - 1
    b = np.ones(n)
    a0 = oracle.AMG(rp, ci, va, max_levels=2, presweeps=1, postsweeps=1, omega=0.8).set_error_scaling(3, 2, 0)
    a1 = oracle.AMG(rp, ci, va, max_levels=2, presweeps=1, postsweeps=1, omega=0.8).set_error_scaling(3, 2, 1)
    x0, x1 = a0.vcycle(b), a1.vcycle(b)
    assert np.array_equal(x0, x1)                       # first cycle computes the scale either way
    A = gallery.to_scipy(rp, ci, va)
    r = b - A @ x0
    y0, y1 = a0.vcycle(r), a1.vcycle(r)                 # second cycle: a1 reuses
    assert not np.array_equal(y0, y1)
    z0, z1 = a0.vcycle(r), a1.vcycle(r)                 # third: a1 recomputes -> same as a0 again
    assert np.array_equal(z0, y0) and np.array_equal(z1, y0)


def test_cg_cycles_and_error_scaling_accelerate_pcg(oracle):
    rp, ci, va = gallery.poisson7pt(18)
    n = rp.shape[0] - 1
    its = {}
    for name, f in {"V": lambda a: a, "CG": lambda a: a.set_cycle("CG"), "CGF": lambda a: a.set_cycle("CGF"),
                    "ES3": lambda a: a.set_error_scaling(3)}.items():
        amg = f(oracle.AMG(rp, ci, va, max_levels=50, presweeps=1, postsweeps=1, omega=0.8))
        # a CG cycle is not a fixed linear operator: flexible PCG around it
        if name.startswith("CG"):
            x, it, hist, conv = oracle.krylov("PCGF", rp, ci, va, np.ones(n), amg=amg, tol=1e-8, max_iters=100)
        else:
            x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=amg, tol=1e-8, max_iters=100)
        assert conv
        its[name] = it
    assert its["CG"] < its["V"] and its["CGF"] < its["V"] and its["ES3"] < its["V"]


def test_structure_reuse_keeps_aggregates_and_recomputes_values(oracle):
    """resetup with structure_reuse_levels = k: coarsenings 1 .. k-1 keep their aggregates, the Galerkin values follow the new matrix"""
    rp, ci, va = gallery.poisson7pt(10, 9, 8)
    n = rp.shape[0] - 1
    rng = np.random.default_rng(9)
    A = gallery.to_scipy(rp, ci, va)
    # a symmetric positive perturbation of the coefficients that changes the strongest-neighbour choices of SIZE_2
    D = 1.0 + 3.0 * rng.random(n)
    import scipy.sparse as sp
    B = (sp.diags(D) @ A @ sp.diags(D)).tocsr()
    B.sort_indices()
    A.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    rp, ci, va, vb = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy(), B.data.copy()
    kw = dict(max_levels=50, presweeps=1, postsweeps=1, omega=0.8)
    first = oracle.AMG(rp, ci, va, **kw)
    fresh = oracle.AMG(rp, ci, vb, **kw)
    assert not np.array_equal(first.level(0)["aggregates"], fresh.level(0)["aggregates"])
    for k in (0, 1, 2, 3, -1):
        re = oracle.AMG(rp, ci, vb, reuse_from=first, structure_reuse_levels=k, **kw)
        kept = re.num_levels() - 1 if k == -1 else max(0, min(k - 1, re.num_levels() - 1))
        for l in range(kept):
            assert np.array_equal(re.level(l)["aggregates"], first.level(l)["aggregates"])
            # values are the Galerkin product of the NEW fine matrix with the OLD aggregates
            Ll, Ln = re.level(l), re.level(l + 1)
            Al = gallery.to_scipy(Ll["row_ptr"], Ll["col_idx"], Ll["values"])
            P = sp.csr_matrix((np.ones(Ll["n"]), (np.arange(Ll["n"]), Ll["aggregates"])), shape=(Ll["n"], Ll["n_coarse"]))
            Ac = (P.T @ Al @ P).toarray()
            got = gallery.to_scipy(Ln["row_ptr"], Ln["col_idx"], Ln["values"]).toarray()
            assert np.allclose(got, Ac, rtol=1e-13, atol=1e-13)
        if kept == 0:
            assert np.array_equal(re.level(0)["aggregates"], fresh.level(0)["aggregates"])
        elif kept < re.num_levels() - 1:
            # the first rebuilt coarsening is SIZE_2 on the (new) matrix of that level, not the old aggregates
            ref = oracle.size2_aggregates(re.level(kept)["row_ptr"], re.level(kept)["col_idx"], re.level(kept)["values"])[0]
            assert np.array_equal(re.level(kept)["aggregates"], ref)
        x, it, hist, conv = oracle.pcg(rp, ci, vb, np.ones(n), amg=re, tol=1e-8, max_iters=100)
        assert conv
