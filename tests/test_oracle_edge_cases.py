"""CPU: degenerate inputs through the oracle (the engine's GPU parity tests use the same shapes): a 1 x 1 system, a diagonal matrix
(no edges: every aggregate is a singleton, coarsening stalls on the first level, src/amg.cu:365-389), a zero right-hand side, a
matrix with an empty row block structure that forces size-1 and size-3 aggregates."""
import numpy as np
import scipy.sparse as sp

from amgx_b200 import gallery


def test_one_by_one_system(oracle):
    rp, ci, va = np.array([0, 1], np.int32), np.array([0], np.int32), np.array([4.0])
    x, it, hist, conv = oracle.pcg(rp, ci, va, np.array([2.0]), tol=1e-10, max_iters=5)
    assert conv and it == 1 and x[0] == 0.5 and hist[-1] == 0.0
    assert oracle.AMG(rp, ci, va, max_levels=10).num_levels() == 1
    assert oracle.ClassicalAMG(rp, ci, va, max_levels=10, interpolator="D2").num_levels() == 1


def test_diagonal_matrix_does_not_coarsen(oracle):
    n = 50
    D = sp.diags(np.arange(1, n + 1, dtype=float)).tocsr()
    rp, ci, va = D.indptr.astype(np.int32), D.indices.astype(np.int32), D.data
    a = oracle.AMG(rp, ci, va, max_levels=10, presweeps=1, postsweeps=1, omega=1.0)
    assert a.num_levels() == 1
    x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=a, tol=1e-10, max_iters=20)
    assert conv and it == 1 and np.allclose(x, 1 / np.arange(1, n + 1), rtol=1e-15)
    c = oracle.ClassicalAMG(rp, ci, va, max_levels=10, interpolator="D2")
    assert c.num_levels() == 1
    for interp in ("D1", "D2", "MULTIPASS"):
        assert oracle.ClassicalAMG(rp, ci, va, max_levels=10, interpolator=interp).num_levels() == 1


def test_zero_right_hand_side_converges_immediately(oracle):
    rp, ci, va = gallery.poisson7pt(5)
    n = rp.shape[0] - 1
    a = oracle.AMG(rp, ci, va, max_levels=10)
    x, it, hist, conv = oracle.pcg(rp, ci, va, np.zeros(n), amg=a, tol=1e-8, max_iters=10)
    assert conv and it == 0 and not x.any()
    x, it, hist, conv = oracle.fgmres(rp, ci, va, np.zeros(n), amg=a, tol=1e-8, max_iters=10, restart=5)
    assert conv and it == 0 and not x.any()


def test_path_graph_gives_odd_sized_aggregates(oracle):
    """a 1-D chain of 7 rows: pairs by handshake, the leftover joins a neighbour's aggregate (size 3), ids ordered by first member"""
    n = 7
    A = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    agg, nagg = oracle.size2_aggregates(rp, ci, va)[:2]
    sizes = np.bincount(agg)
    assert sizes.sum() == n and sizes.min() >= 1 and sizes.max() <= 3 and nagg == len(sizes)
    first = [int(np.nonzero(agg == a)[0][0]) for a in range(nagg)]
    assert first == sorted(first)
    for a in range(nagg):               # aggregates are connected pieces of the chain
        m = np.nonzero(agg == a)[0]
        assert m.max() - m.min() == len(m) - 1


def missing_diagonal_matrix(n=200, seed=1):
    """the generator of the reference's zero_in_diagonal_handling test (src/tests/zero_in_diagonal_handling.cu:57-118): unit values,
    up to 10 random columns per row, five rows have no diagonal entry at all"""
    rng = np.random.default_rng(seed)
    rows, cols, left = [], [], 5
    for i in range(n):
        c = set()
        if left > 0 and rng.integers(5) != 0:
            left -= 1
        else:
            c.add(i)
        for _ in range(max(1, int(rng.random() * (10 - len(c))))):
            c.add(int(rng.integers(n)))
        for j in sorted(c):
            rows.append(i)
            cols.append(j)
    A = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()


def test_structurally_missing_diagonal_entries_give_no_nan(oracle):
    """ImplicitZeroInDiagonal: selectors, coarse generators, colourings and smoothers must survive rows without a diagonal entry
    (the guards: isNotCloseToZero(d) ? d : epsilon(d), block_jacobi_solver.cu:29-50; Einv of DILU: res != 0 ? 1 / res : res)"""
    rp, ci, va = missing_diagonal_matrix()
    n = rp.shape[0] - 1
    for sm in ("BLOCK_JACOBI", "JACOBI_L1", "MULTICOLOR_DILU", "MULTICOLOR_GS"):
        a = oracle.AMG(rp, ci, va, max_levels=10, presweeps=1, postsweeps=1, omega=0.8, smoother=sm)
        assert a.num_levels() >= 2
        for l in range(a.num_levels()):
            assert np.isfinite(a.level(l)["values"]).all()
        x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=a, tol=1e-8, max_iters=3)
        assert np.isfinite(x).all() and np.isfinite(hist).all(), sm
    for interp in ("D1", "D2", "MULTIPASS"):
        c = oracle.ClassicalAMG(rp, ci, va, max_levels=10, interpolator=interp)
        x, it, hist, conv = oracle.fgmres(rp, ci, va, np.ones(n), amg=c, tol=1e-8, max_iters=3, restart=3)
        assert np.isfinite(x).all() and np.isfinite(hist).all(), interp


def poisson27(nx, ny, nz):
    """27-point stencil, centre 26, the 26 neighbours -1 (cusp::gallery::poisson27pt, what generatePoissonForTest(..., 27, ...) builds)"""
    idx = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    rows, cols, vals = [], [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                src = idx[max(0, -dz):nz - max(0, dz), max(0, -dy):ny - max(0, dy), max(0, -dx):nx - max(0, dx)]
                dst = idx[max(0, dz):nz - max(0, -dz), max(0, dy):ny - max(0, -dy), max(0, dx):nx - max(0, -dx)]
                rows.append(src.ravel())
                cols.append(dst.ravel())
                vals.append(np.full(src.size, 26.0 if (dx, dy, dz) == (0, 0, 0) else -1.0))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(idx.size, idx.size))
    A.sort_indices()
    return A


def test_aggregates_coarsening_factor(oracle):
    """the reference's AggregatesCoarseningFactor unit test (src/tests/aggregates_coarsening_factor.cu): on a 20^3 27-point Poisson matrix
    with values perturbed by up to 1/50, max_unassigned_percentage = 0.1, deterministic: SIZE_2 leaves < 1.1 / 2 of the rows,
    SIZE_4 < 1.1 / 4; every row gets a valid aggregate id"""
    A = poisson27(20, 20, 20)
    A.data += np.random.default_rng(30).random(A.nnz) / 50.0
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    n = rp.shape[0] - 1
    for fn, bound in ((oracle.size2_aggregates, 1.1 / 2), (oracle.size4_aggregates, 1.1 / 4)):
        agg, nagg = fn(rp, ci, va, max_unassigned=0.1)[:2]
        assert nagg / n < bound, (fn.__name__, nagg / n)
        assert agg.min() == 0 and agg.max() == nagg - 1 and np.unique(agg).shape[0] == nagg


def poisson2d(points, nx, ny):
    """cusp::gallery::poisson5pt / poisson9pt: centre 4 (8), neighbours -1"""
    idx = np.arange(nx * ny).reshape(ny, nx)
    rows, cols, vals = [], [], []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if points == 5 and dx != 0 and dy != 0:
                continue
            src = idx[max(0, -dy):ny - max(0, dy), max(0, -dx):nx - max(0, dx)]
            dst = idx[max(0, dy):ny - max(0, -dy), max(0, dx):nx - max(0, -dx)]
            rows.append(src.ravel())
            cols.append(dst.ravel())
            vals.append(np.full(src.size, float(points - 1) if (dx, dy) == (0, 0) else -1.0))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(idx.size, idx.size))
    A.sort_indices()
    return A


def test_fgmres_convergence_poisson(oracle):
    """the reference's FGMRESConvergencePoisson unit test (src/tests/fgmres_convergence_poisson.cu): FGMRES with a full-length restart and
    one BLOCK_JACOBI sweep as preconditioner on 5- and 9-point Poisson grids 5x5 .. 10x10, tolerance 1e-14: true relative residual < 1e-5"""
    for size in range(5, 11):
        for points in (5, 9):
            A = poisson2d(points, size, size)
            rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
            n = rp.shape[0] - 1
            b = np.ones(n)
            x, it, hist, conv = oracle.fgmres(rp, ci, va, b, jacobi_omega=0.9, tol=1e-14, max_iters=n, restart=n)
            assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) < 1e-5, (points, size)


def test_fgmres_zero_initial_residual(oracle):
    """the reference's FGMRESZeroInitialResidual unit test: A = diag(2, 4), b = 0, x = 0 -- convergence is reported without an iteration"""
    rp, ci, va = np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), np.array([2.0, 4.0])
    x, it, hist, conv = oracle.fgmres(rp, ci, va, np.zeros(2), jacobi_omega=0.9, tol=1e-14, max_iters=2, restart=2)
    assert conv and it == 0 and not x.any() and np.isfinite(hist).all()


def test_scalar_smoothers_poisson(oracle):
    """the reference's ScalarSmootherPoisson unit test (src/tests/scalar_smoother_poisson.cu): 1000 sweeps of BLOCK_JACOBI, MULTICOLOR_DILU
    and symmetric MULTICOLOR_GS (weight 1, MIN_MAX colours, nothing left uncoloured) on the 9-point 10 x 10 Poisson matrix bring the
    residual of b = 1, x0 = 0 below 1e-5"""
    A = poisson2d(9, 10, 10)
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    n = rp.shape[0] - 1
    b = np.ones(n)
    oracle.set_uncolored_fraction(0.0)
    try:
        for sm in ("BLOCK_JACOBI", "MULTICOLOR_DILU", "MULTICOLOR_GS"):
            a = oracle.AMG(rp, ci, va, max_levels=1, coarsest_sweeps=1000, smoother=sm, omega=1.0)     # one level: the cycle IS the smoother
            if sm == "MULTICOLOR_GS":
                a.set_symmetric_gs(True)
            x, it, hist, conv = oracle.amg_solve(a, b, tol=1e-300, max_iters=1)
            r = np.linalg.norm(b - A @ x)
            assert r < np.linalg.norm(b) and r < 1e-5, (sm, r)
    finally:
        oracle.set_uncolored_fraction(0.15)


def test_dense_lu_solve_poisson3d(oracle):
    """the reference's DenseLUSolverTest_Solve_Poisson3D unit test (src/tests/dense_lu.cu:215-258; 27-point Poisson, b = 1, residual norm
    < 1e-12) on the restatement of the coarse solver; 10^3 rows here, the opt-in GPU test runs the reference's 16^3"""
    A = poisson27(10, 10, 10)
    n = A.shape[0]
    x, lu, ipiv = oracle.dense_lu_solve(A.toarray(), np.ones(n))
    assert np.linalg.norm(np.ones(n) - A @ x) < 1e-12


def test_valid_coloring_on_random_symmetric_structures(oracle):
    """the reference's MinMaxColoringTest unit test (src/tests/valid_coloring.cu), coloring_level 1, max_uncolored_percentage 0: on random
    structurally symmetric matrices no two neighbouring coloured rows share a colour -- for MIN_MAX and for PARALLEL_GREEDY"""
    rng = np.random.default_rng(10125)
    for trial in range(10):
        n = int(rng.integers(1, 10000))
        per_row = max(int(rng.integers(0, 10)), 1)
        r = np.repeat(np.arange(n), per_row)
        c = rng.integers(0, n, r.shape[0])
        S = sp.coo_matrix((np.ones(r.shape[0]), (r, c)), shape=(n, n)).tocsr()
        S = (S + S.T + sp.identity(n)).tocsr()
        S.sort_indices()
        rp, ci = S.indptr.astype(np.int32), S.indices.astype(np.int32)
        rows = np.repeat(np.arange(n), np.diff(rp))
        off = rows != ci
        for fn in (oracle.color_min_max, oracle.color_parallel_greedy):
            nc, colors, srows, offs = fn(rp, ci, 0.0)
            both = off & (colors[rows] != 0) & (colors[ci] != 0)
            assert not np.any(colors[rows][both] == colors[ci][both]), (fn.__name__, trial)
            assert np.array_equal(np.sort(srows), np.arange(n)) and offs[0] == 0 and offs[-1] == n


def random_unsymmetric_with_a_zero(n=100, seed=31):
    """generateMatrixRandomStruct + random_fill + random_add_zeros of the reference's zero_values_handling test: random, structurally
    unsymmetric rows (diagonal included), positive random values, the first off-diagonal value of one row replaced by an explicit zero"""
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for i in range(n):
        c = {i}
        for _ in range(int(rng.integers(1, 10))):
            c.add(int(rng.integers(n)))
        for j in sorted(c):
            rows.append(i)
            cols.append(j)
    A = sp.csr_matrix((rng.random(len(rows)) + 0.1, (rows, cols)), shape=(n, n))
    A.sort_indices()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    while True:
        r = int(rng.integers(n - 1))
        if ci[rp[r]] != r:
            va[rp[r]] = 0.0
            return rp, ci, va


def test_explicit_zero_values_and_unsymmetric_structure_give_no_nan(oracle):
    """ExplicitZeroValues (src/tests/zero_values_handling.cu): every selector, coarse generator, colouring, smoother, strength,
    classical selector and interpolator must get through such a matrix without NaNs or errors"""
    for seed in (31, 32, 33):
        rp, ci, va = random_unsymmetric_with_a_zero(seed=seed)
        n = rp.shape[0] - 1
        for sel in ("SIZE_2", "SIZE_4"):
            for sm in ("BLOCK_JACOBI", "JACOBI_L1", "MULTICOLOR_DILU", "MULTICOLOR_GS"):
                a = oracle.AMG(rp, ci, va, max_levels=10, min_coarse_rows=2, smoother=sm, selector=sel)
                x, it, hist, conv = oracle.pcg(rp, ci, va, np.ones(n), amg=a, tol=1e-8, max_iters=2)
                assert np.isfinite(x).all() and np.isfinite(hist).all(), (seed, sel, sm)
        for s in ("PMIS", "HMIS"):
            for interp in ("D1", "D2", "MULTIPASS"):
                c = oracle.ClassicalAMG(rp, ci, va, max_levels=10, interpolator=interp, selector=s)
                x, it, hist, conv = oracle.fgmres(rp, ci, va, np.ones(n), amg=c, tol=1e-8, max_iters=2, restart=2)
                assert np.isfinite(x).all() and np.isfinite(hist).all(), (seed, s, interp)


def test_norms_against_numpy(oracle):
    """norm_tests.cu of the reference: L1 / L2 / LMAX against host values within type-epsilon x size; here the oracle's norm functions
    (the checker of every residual history) against numpy, and through the PCG monitor: history entry 0 is the norm of b"""
    import ctypes as C
    rng = np.random.default_rng(5)
    lib = oracle.lib()
    for n in (1, 7, 1000, 65537):
        v = np.ascontiguousarray(rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n))
        p = v.ctypes.data_as(C.c_void_p)
        eps = np.finfo(np.float64).eps * n
        assert abs(lib.orc_nrm1(n, p) - np.abs(v).sum()) <= eps * np.abs(v).sum()
        assert abs(lib.orc_nrm2(n, p) - np.linalg.norm(v)) <= eps * np.linalg.norm(v)
        assert lib.orc_nrmmax(n, p) == np.abs(v).max()
    from amgx_b200 import gallery
    rp, ci, va = gallery.poisson7pt(6)
    b = rng.standard_normal(rp.shape[0] - 1)
    for norm, f in (("L1", lambda r: np.abs(r).sum()), ("L2", np.linalg.norm), ("LMAX", lambda r: np.abs(r).max())):
        x, it, hist, conv = oracle.pcg(rp, ci, va, b, jacobi_omega=0.8, tol=1e-8, max_iters=100, norm=norm)
        assert conv and abs(hist[0] - f(b)) <= 1e-14 * f(b)
        r = b - gallery.to_scipy(rp, ci, va) @ x
        assert f(r) <= 1.0001e-8 * f(b) + 1e-15
