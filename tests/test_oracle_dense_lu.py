"""CPU: the DENSE_LU_SOLVER restatement (oracle) -- LU with partial pivoting in the engine's operation order -- against
numpy, and the AMG V-cycle with a direct coarsest-level solve."""
import numpy as np
import pytest
import scipy.sparse as sp

from amgx_b200 import gallery


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (64, 3), (129, 4)])
def test_dense_lu_matches_numpy(oracle, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + np.diag(rng.standard_normal(n) * 0.1)     # pivoting is exercised: no dominance
    b = rng.standard_normal(n)
    x, lu, ipiv = oracle.dense_lu_solve(A, b)
    assert np.allclose(A @ x, b, rtol=0, atol=1e-10 * max(1.0, np.abs(b).max()) * np.linalg.cond(A))
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9 * np.linalg.cond(A), atol=1e-12)
    # P A = L U with the recorded interchanges
    P = np.arange(n)
    for k in range(n):
        P[[k, ipiv[k]]] = P[[ipiv[k], k]]
    L = np.tril(lu, -1) + np.eye(n)
    U = np.triu(lu)
    assert np.allclose(L @ U, A[P], rtol=0, atol=1e-12 * np.abs(A).max() * n)
    assert np.all(np.abs(np.tril(lu, -1)) <= 1.0 + 1e-15)       # partial pivoting bounds the multipliers


def test_dense_lu_first_maximum_is_the_pivot(oracle):
    A = np.array([[1.0, 2.0, 0.0], [-3.0, 1.0, 1.0], [3.0, 0.0, 1.0]])     # |a10| == |a20|: idamax takes the first
    x, lu, ipiv = oracle.dense_lu_solve(A, np.ones(3))
    assert ipiv[0] == 1
    assert np.allclose(A @ x, np.ones(3))


@pytest.mark.parametrize("algo", ["aggregation", "classical"])
def test_amg_with_dense_lu_coarse_solver(oracle, algo):
    rp, ci, va = gallery.poisson7pt(14)
    n = rp.shape[0] - 1
    kw = dict(coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=64)
    if algo == "aggregation":
        amg = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.8, **kw)
        ref = oracle.AMG(rp, ci, va, max_levels=50, presweeps=0, postsweeps=3, omega=0.8)
    else:
        c = dict(max_levels=50, presweeps=2, postsweeps=2, smoother="JACOBI_L1", omega=1.0, strength_threshold=0.25, max_row_sum=0.9,
                 interpolator="D2", aggressive_levels=1, interp_max_elements=4)
        amg = oracle.ClassicalAMG(rp, ci, va, **c, **kw)
        ref = oracle.ClassicalAMG(rp, ci, va, **c)
    # dense_lu_num_rows becomes min_coarse_rows (src/amg.cu:1154-1157): coarsening stops as soon as a level has <= 64 rows or
    # the next one would have fewer than 64 -- fewer levels, a coarsest level of the order of 64 rows
    assert amg.num_levels() < ref.num_levels()
    last = amg.level(amg.num_levels() - 1)
    assert last["n"] <= 4 * 64 and amg.level(amg.num_levels() - 2)["n"] > 64
    b = np.ones(n)
    x1, it1, h1, c1 = oracle.pcg(rp, ci, va, b, amg=amg, tol=1e-8, max_iters=100)
    x2, it2, h2, c2 = oracle.pcg(rp, ci, va, b, amg=ref, tol=1e-8, max_iters=100)
    assert c1 and c2 and it1 <= it2 + 1
    A = gallery.to_scipy(rp, ci, va)
    assert np.linalg.norm(b - A @ x1) <= 1.01e-8 * np.linalg.norm(b)
    # a single-level hierarchy with the direct solve is an exact solver
    rp2, ci2, va2 = gallery.poisson7pt(4)
    one = oracle.AMG(rp2, ci2, va2, max_levels=1, coarse_solver="DENSE_LU_SOLVER", dense_lu_num_rows=128)
    y = one.vcycle(np.ones(64))
    assert np.allclose(gallery.to_scipy(rp2, ci2, va2) @ y, np.ones(64), atol=1e-12)
