"""GPU: the drop-in Python API (matrix / objectivefunc / optimizer) against dense NumPy equivalents
and the reference's golden outputs.  Test cases follow the reference's own suites
(/root/reference/test/test_matrix.py, test_objectivefunc.py, test_optimizer.py)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api(build_lib):
    assert torch.cuda.is_available()
    import admmsolver_b200.matrix as M
    import admmsolver_b200.objectivefunc as F
    import admmsolver_b200.optimizer as O
    return M, F, O


def _rc(rs, *shape):
    return rs.randn(*shape) + 1j * rs.randn(*shape)


# ------------------------------------------------------------------ matrix.py
def test_matmul_all_pairs(api):
    """test_matrix.py:11-57."""
    M, F, O = api
    rs = np.random.RandomState(100)
    n1, n2, n3 = 12, 12, 4
    left = [M.DiagonalMatrix(np.ones(n1)), M.ScaledIdentityMatrix(n1, 1 + 1j),
            M.PartialDiagonalMatrix(_rc(rs, 3, 3), rest_dims=(4,)), M.DenseMatrix(_rc(rs, n1, n2))]
    right = [M.DenseMatrix(_rc(rs, n2, n3)), M.ScaledIdentityMatrix((n2, n3), 1 + 1j),
             M.PartialDiagonalMatrix(_rc(rs, 3, 1), rest_dims=(4,))]
    for l in left:
        for r in right:
            lr = l @ r
            assert isinstance(lr, M.MatrixBase)
            np.testing.assert_allclose(lr.asmatrix(), l.asmatrix() @ r.asmatrix(), atol=1e-13)
    n1, n2, n3 = 4, 12, 12
    left = [M.DenseMatrix(_rc(rs, n1, n2)), M.PartialDiagonalMatrix(_rc(rs, 1, 3), rest_dims=(4,))]
    right = [M.DiagonalMatrix(np.ones(n3)), M.ScaledIdentityMatrix(n3, 1 + 1j),
             M.PartialDiagonalMatrix(_rc(rs, 3, 3), rest_dims=(4,)), M.DenseMatrix(_rc(rs, n2, n3))]
    for l in left:
        for r in right:
            lr = l @ r
            assert isinstance(lr, M.MatrixBase)
            np.testing.assert_allclose(lr.asmatrix(), l.asmatrix() @ r.asmatrix(), atol=1e-13)


def test_mul_transpose_conj_add_inv(api):
    """test_matrix.py:60-166."""
    M, F, O = api
    rs = np.random.RandomState(100)
    n = 4
    mats = [M.DiagonalMatrix(_rc(rs, n)), M.ScaledIdentityMatrix(n, 1 + 1j), M.PartialDiagonalMatrix(_rc(rs, 2, 2), (2,)),
            M.DenseMatrix(_rc(rs, n, n)), M.DiagonalMatrix(_rc(rs, 2), shape=(4, 2)), M.ScaledIdentityMatrix((2, 4), 2.0)]
    for m in mats:
        d = m.asmatrix()
        np.testing.assert_allclose((m * (2.0 + 1j)).asmatrix(), d * (2.0 + 1j), atol=1e-14)
        np.testing.assert_allclose((3.0 * m).asmatrix(), 3.0 * d, atol=1e-14)
        np.testing.assert_allclose(m.T.asmatrix(), d.T, atol=1e-14)
        np.testing.assert_allclose(m.conj().asmatrix(), d.conj(), atol=1e-14)
        np.testing.assert_allclose((-m).asmatrix(), -d, atol=1e-14)
    sq = mats[:4]
    for a in sq:
        for b in sq:
            np.testing.assert_allclose((a + b).asmatrix(), a.asmatrix() + b.asmatrix(), atol=1e-13)
            np.testing.assert_allclose((a - b).asmatrix(), a.asmatrix() - b.asmatrix(), atol=1e-13)
    for m in sq:
        inv_m = m.inv()
        assert isinstance(inv_m, M.MatrixBase)
        np.testing.assert_allclose(inv_m.asmatrix() @ m.asmatrix(), np.identity(n), rtol=0, atol=1e-12)
    with pytest.raises(RuntimeError):
        M.ScaledIdentityMatrix((2, 4), 2.0).inv()
    with pytest.raises(AssertionError):
        M.ScaledIdentityMatrix(3, 1)               # int coefficients are rejected (matrix.py:134-135)


def test_structure_preservation(api):
    """test_matrix.py:110-150: Diagonal (+|@) PartialDiagonal stays PartialDiagonal when constant along rest."""
    M, F, O = api
    rs = np.random.RandomState(1)
    a = M.DiagonalMatrix(np.repeat(rs.randn(3), 4))
    b = M.PartialDiagonalMatrix(rs.randn(3, 3), (4,))
    for r in (a + b, a @ b, b + b, b @ b):
        assert isinstance(r, M.PartialDiagonalMatrix)
    np.testing.assert_allclose((a + b).asmatrix(), a.asmatrix() + b.asmatrix(), atol=1e-14)
    np.testing.assert_allclose((a @ b).asmatrix(), a.asmatrix() @ b.asmatrix(), atol=1e-14)
    c = M.DiagonalMatrix(rs.randn(12))
    assert isinstance(c + b, M.DenseMatrix)
    np.testing.assert_allclose((c + b).asmatrix(), c.asmatrix() + b.asmatrix(), atol=1e-14)


@pytest.mark.parametrize("n,m", [(4, 4), (2, 4), (4, 2)])
def test_matvec_and_batched(api, n, m):
    """test_matrix.py:169-233 incl. rectangular and (m, nbatch) right-hand sides."""
    M, F, O = api
    rs = np.random.RandomState(100)
    mats = [M.DiagonalMatrix(np.ones(min(n, m)), shape=(n, m)), M.ScaledIdentityMatrix((n, m), 1 + 1j),
            M.PartialDiagonalMatrix(_rc(rs, n // 2, m // 2), (2,)),
            M.PartialDiagonalMatrix(M.DiagonalMatrix(_rc(rs, min(n // 2, m // 2)), (n // 2, m // 2)), (2,)),
            M.DenseMatrix(_rc(rs, n, m))]
    if n == m:
        mats.append(M.PartialDiagonalMatrix(M.ScaledIdentityMatrix(n // 2, 1.0), (2,)))
    for v in (np.ones(m), _rc(rs, m, 3)):
        for mat in mats:
            mv = mat @ v
            assert isinstance(mv, np.ndarray)
            np.testing.assert_allclose(mv, mat.asmatrix() @ v, atol=1e-14)
            dv = mat @ torch.from_numpy(np.asarray(v, dtype=complex)).cuda()
            assert isinstance(dv, torch.Tensor) and dv.is_cuda
            np.testing.assert_allclose(dv.cpu().numpy(), mat.asmatrix() @ v, atol=1e-14)


def test_rectangular_diagonal_product_and_helpers(api):
    """test_matrix.py:236-257."""
    M, F, O = api
    rs = np.random.RandomState(100)
    a = M.DiagonalMatrix(rs.randn(2), shape=(4, 2))
    b = M.DiagonalMatrix(rs.randn(2), shape=(2, 4))
    ab = a @ b
    ref = np.zeros(4)
    ref[:2] = a.diagonals * b.diagonals
    np.testing.assert_allclose(ab.diagonals, ref)
    np.testing.assert_allclose(M._vecprod(np.ones(1), np.ones(2), 3), [1, 0, 0])
    np.testing.assert_allclose(M._pad_by_zero(np.ones(1), 3), [1, 0, 0])
    assert M.matrix_hash(M.DenseMatrix(np.eye(2))) == M.matrix_hash(np.eye(2))
    assert M.identity(3).coeff == 1.0 and isinstance(M.asmatrixtype(np.eye(2)), M.DenseMatrix)


# ------------------------------------------------------------------ objectivefunc.py
def test_terms_golden(api):
    """test_objectivefunc.py:34-144 cases; expected values from the reference itself."""
    M, F, O = api
    g = golden("terms")
    x = F.LeastSquares(2.0, g["A"], g["y"]).solve(g["h"], M.DenseMatrix(g["mu"]))
    assert rel(x, g["x_ls"]) < TOL
    cls = F.ConstrainedLeastSquares(2.0, g["A"], g["y"], g["C"], g["D"])
    xc = cls.solve(g["h"], M.DenseMatrix(g["mu"]))
    assert rel(xc, g["x_cls"]) < TOL
    assert np.abs(g["C"] @ xc - g["D"]).max() < 1e-10             # test_objectivefunc.py:101
    xp = F.LeastSquares(0.3, M.PartialDiagonalMatrix(g["a2"], (20,)), g["y2"]).solve(g["h2"], M.ScaledIdentityMatrix(20, 1.5))
    assert rel(xp, g["x_partial"]) < TOL
    xl = F.L1Regularizer(0.3, 7).solve(g["hl"], M.DiagonalMatrix(np.linspace(0.5, 2.0, 7)))
    assert xl.dtype == np.float64 and np.array_equal(xl == 0, g["x_l1"] == 0) and rel(xl, g["x_l1"]) < 1e-14
    xn = F.NonNegativePenalty(7).solve(g["hl"] + 0.1j, M.ScaledIdentityMatrix(7, 0.7))
    assert rel(xn, g["x_nn"]) < 1e-14 and (xn >= 0).all()
    with pytest.raises(AssertionError):
        F.L1Regularizer(0.3, 7).solve(g["hl"], M.DenseMatrix(np.eye(7)))
    # objective values
    ls = F.LeastSquares(2.0, g["A"], g["y"])
    assert abs(ls(x) - 2.0 * np.linalg.norm(g["y"] - g["A"] @ x) ** 2) < 1e-12
    assert abs(F.L1Regularizer(0.3, 7)(xl) - 0.3 * np.abs(xl).sum()) < 1e-14


def test_ridge_with_l2(api):
    """test_optimizer.py:85-109 (L2Regularizer)."""
    M, F, O = api
    rs = np.random.RandomState(100)
    y, A, B = _rc(rs, 2), _rc(rs, 2, 2), _rc(rs, 1, 2)
    model = O.Model([F.LeastSquares(1.0, A, y), F.L2Regularizer(1, B)], [(1, 0, M.identity(2), M.identity(2))])
    opt = O.SimpleOptimizer(model)
    opt.solve(niter=100, update_h=True)
    x_ref = np.linalg.inv(A.conj().T @ A + B.conj().T @ B) @ A.conj().T @ y
    np.testing.assert_allclose(opt.x[0], x_ref, atol=np.abs(x_ref).max() * 1e-8)


# ------------------------------------------------------------------ optimizer.py
def test_lasso_and_basis_pursuit_dropin(api):
    """test_optimizer.py:13-82 through the drop-in API (fused pattern A engine)."""
    M, F, O = api
    from admmsolver_b200 import problems
    g = golden("lasso_1x2")
    opt = O.SimpleOptimizer(O.Model([F.LeastSquares(1.0, g["A"], g["y"]), F.L1Regularizer(0.1, 2)],
                                    [(1, 0, M.identity(2), M.identity(2))]))
    assert opt._plan_kind == "bp"
    opt.solve(100)
    assert len(opt._primal_residual) == len(g["primal"])
    for k, name in enumerate(("x0", "x1")):
        assert rel(opt.x[k], g[name]) < TOL
    A, y, xa = problems.basis_pursuit_instance(100, 1000, 20, 1234)
    g = golden("bp_notebook")
    opt = O.SimpleOptimizer(O.Problem([F.LeastSquares(1.0, A, g["y"]), F.L1Regularizer(1e-1, 1000)],
                                      [O.EqualityCondition(1, 0, M.identity(1000), M.identity(1000))]))
    opt.solve(100)
    np.testing.assert_allclose(opt.x[0], xa, atol=1e-2 * np.abs(xa).max(), rtol=0)
    assert opt.x[0].dtype == np.complex128
    assert rel(opt.x[0], g["x0"]) < TOL and rel(opt._h[1, 0], g["h10"]) < 1e-8
    assert abs(opt(opt.x) - g["objective"]) / g["objective"] < TOL
    assert opt._mu[1, 0] == g["mu10"]
    assert rel(opt._primal_residual, g["primal"]) < 1e-8 and rel(opt._dual_residual, g["dual"]) < 1e-8


def test_generic_executor_equals_fused_and_golden(api):
    """The generic device executor (callback forces it) walks the same trajectory as the reference."""
    M, F, O = api
    g = golden("bp_tall")
    mk = lambda: O.SimpleOptimizer(O.Model([F.LeastSquares(0.7, g["A"], g["y"]), F.L1Regularizer(0.05, 40)],
                                           [(1, 0, M.identity(40), M.identity(40))]))
    calls = []
    opt = mk()
    opt.solve(250, callback=lambda: calls.append(1))
    assert len(calls) == 250
    assert rel(opt.x[0], g["x0"]) < TOL and rel(opt.x[1], g["x1"]) < TOL and opt._mu[1, 0] == g["mu10"]
    assert rel(opt._primal_residual, g["primal"]) < 1e-8
    # step-wise public methods
    opt2 = mk()
    with pytest.raises(AttributeError):
        opt2.residual()
    for it in range(3):
        opt2.one_sweep(update_h=True)
        p, d = opt2.residual()
        assert abs(p - g["primal"][it]) / g["primal"][it] < 1e-9
        opt2.check_convergence(1e-12)
        if it % 100 == 0:
            opt2.update_mu()


def test_generic_three_term_dense_couplings(api):
    """Dense rectangular E, rectangular diagonal partner: not a fused pattern -> generic executor."""
    M, F, O = api
    g = golden("generic3")
    lst, l1, nn = F.LeastSquares(1.3, g["A"], g["y"]), F.L1Regularizer(0.2, 5), F.NonNegativePenalty(4)
    conds = [(0, 1, g["E1"], M.identity(5)), (0, 2, g["P"], M.DiagonalMatrix(np.linspace(1.0, 2.0, 4)))]
    opt = O.SimpleOptimizer(O.Model([lst, l1, nn], conds), mu=0.7)
    assert opt._plan_kind is None
    opt.solve(150, interval_update_mu=20)
    for k in range(3):
        assert rel(opt.x[k], g[f"x{k}"]) < 1e-9
    assert opt._mu[1, 0] == g["mu10"] and opt._mu[2, 0] == g["mu20"]
    assert rel(opt._primal_residual, g["primal"]) < 1e-8
    assert abs(opt(opt.x) - g["objective"]) / g["objective"] < 1e-9


def test_generic_run_ahead_early_stop_and_resume(api):
    """The generic executor runs chunks of iterations ahead of the host (one synchronisation per chunk) and rolls
    back when the stopping test fired inside a chunk: same iteration count, history, penalties and state as the
    iteration-by-iteration path (a callback forces that one) -- for an early stop in the middle of a chunk, for a
    chunk cut by niter, and for a second solve() that continues."""
    M, F, O = api
    g = golden("generic3")
    conds = [(0, 1, g["E1"], M.identity(5)), (0, 2, g["P"], M.DiagonalMatrix(np.linspace(1.0, 2.0, 4)))]
    mk = lambda: O.SimpleOptimizer(O.Model([F.LeastSquares(1.3, g["A"], g["y"]), F.L1Regularizer(0.2, 5), F.NonNegativePenalty(4)],
                                           conds), mu=0.7)
    for niter, rtol in ((400, 1e-6), (57, 1e-12), (400, 1e-3)):
        a, b = mk(), mk()
        a.solve(niter, interval_update_mu=20, rtol=rtol)
        b.solve(niter, interval_update_mu=20, rtol=rtol, callback=lambda: None)
        assert len(a._primal_residual) == len(b._primal_residual), (niter, rtol)
        if rtol > 1e-12:
            assert len(a._primal_residual) < niter                   # really an early stop
        assert rel(a._primal_residual, b._primal_residual) < 1e-12 and rel(a._dual_residual, b._dual_residual) < 1e-12
        for k in range(3):
            assert rel(a.x[k], b.x[k]) < 1e-13
        assert a._mu[1, 0] == b._mu[1, 0] and a._mu[2, 0] == b._mu[2, 0]
        a.solve(30, interval_update_mu=20, rtol=1e-12)
        b.solve(30, interval_update_mu=20, rtol=1e-12, callback=lambda: None)
        for k in range(3):
            assert rel(a.x[k], b.x[k]) < 1e-13


def test_spm_notebook_flow_dropin(api):
    """spm.ipynb:243-259 through the drop-in API: single problem (fused pattern B engine)."""
    M, F, O = api
    g = golden("spm_small")
    L, Nw = g["s"].size, g["P"].shape[0]
    lstsq = F.ConstrainedLeastSquares(1.0, -M.DiagonalMatrix(g["s"]), g["g"], g["C"], np.array([1]))
    conds = [(0, 1, M.identity(L), M.identity(L)), (0, 2, g["P"], M.identity(Nw))]
    opt = O.SimpleOptimizer(O.Problem([lstsq, F.L1Regularizer(float(g["lam"]), L), F.NonNegativePenalty(Nw)], conds), mu=0.1)
    assert opt._plan_kind == "spm"
    opt.solve(700)
    for k in range(3):
        assert rel(opt.x[k], g[f"x{k}"]) < TOL
    assert rel(opt._h[2, 0], g["h20"]) < 1e-8
    assert abs((g["C"] @ opt.x[0])[0] - 1) < 1e-12                 # spm.ipynb:270
    assert abs(opt(opt.x) - g["objective"]) / g["objective"] < TOL
    assert len(opt._primal_residual) == 700 and rel(opt._primal_residual, g["primal"]) < 1e-8


def test_spm_packed_partialdiagonal_dropin(api):
    """The reference's own batching: PartialDiagonalMatrix-packed operators -> fused engine, batch-wide mu."""
    M, F, O = api
    g = golden("spm_packed")
    nb, L, Nw = 6, g["s"].size, g["P"].shape[0]
    rest = (nb,)
    lstsq = F.ConstrainedLeastSquares(1.0, M.PartialDiagonalMatrix(-M.DiagonalMatrix(g["s"]), rest), g["g"].ravel(),
                                      M.PartialDiagonalMatrix(g["C"], rest), np.ones(nb))
    conds = [(0, 1, M.identity(L * nb), M.identity(L * nb)),
             (0, 2, M.PartialDiagonalMatrix(g["P"], rest), M.identity(Nw * nb))]
    opt = O.SimpleOptimizer(O.Model([lstsq, F.L1Regularizer(float(g["lam"]), L * nb), F.NonNegativePenalty(Nw * nb)], conds),
                            mu=float(g["mu"]))
    assert opt._plan_kind == "spm"
    opt.solve(250)
    opt.solve(150)                                                 # resume like the reference
    # two calls restart the `iter % interval` phase: compare with the oracle doing the same
    from oracle import flat
    st = flat.spm_solve(g["s"], g["P"], g["C"], np.ones(nb), g["g"], float(g["lam"]), 250, mu=float(g["mu"]))
    st = flat.spm_solve(g["s"], g["P"], g["C"], np.ones(nb), g["g"], float(g["lam"]), 150, mu=float(g["mu"]), state=st)
    assert rel(opt.x[0], st.x0.ravel()) < TOL and rel(opt.x[2], st.x2.ravel()) < TOL
    assert opt._mu[1, 0] == st.mu10 and opt._mu[2, 0] == st.mu20
    # and the generic executor on the same packed model (callback) follows the reference golden
    opt2 = O.SimpleOptimizer(O.Model([lstsq, F.L1Regularizer(float(g["lam"]), L * nb), F.NonNegativePenalty(Nw * nb)], conds),
                             mu=float(g["mu"]))
    opt2.solve(30, callback=lambda: None)
    assert rel(opt2._primal_residual, g["primal"][:30]) < 1e-8


def test_spm_packed_dropin_folded_pass(api):
    """The same packed model with a sampling matrix that has the exact parity of the IR basis: the drop-in optimizer's
    plan takes the folded pass (admm_spm_dims.fold); solve / resume / warm start from the exported x and h follow the
    oracle, and the state a caller reads between the calls (x[2], h of pair (2,0)) is the unfolded one."""
    M, F, O = api
    from admmsolver_b200 import problems
    from oracle import flat
    g = golden("spm_packed")
    nb, L, Nw = 6, g["s"].size, g["P"].shape[0]
    P = problems.symmetrize_sampling(np.ascontiguousarray(g["P"]))
    assert Nw % 2 == 0 and np.abs(P - g["P"]).max() < 1e-6 * np.abs(P).max()
    rest = (nb,)

    def model():
        lstsq = F.ConstrainedLeastSquares(1.0, M.PartialDiagonalMatrix(-M.DiagonalMatrix(g["s"]), rest), g["g"].ravel(),
                                          M.PartialDiagonalMatrix(g["C"], rest), np.ones(nb))
        conds = [(0, 1, M.identity(L * nb), M.identity(L * nb)),
                 (0, 2, M.PartialDiagonalMatrix(P, rest), M.identity(Nw * nb))]
        return O.Model([lstsq, F.L1Regularizer(float(g["lam"]), L * nb), F.NonNegativePenalty(Nw * nb)], conds)

    opt = O.SimpleOptimizer(model(), mu=float(g["mu"]))
    assert opt._plan_kind == "spm"
    opt.solve(250)
    assert opt._plan.eng.fold
    x_mid = [np.array(v) for v in opt.x]
    opt.solve(150)
    st1 = flat.spm_solve(g["s"], P, g["C"], np.ones(nb), g["g"], float(g["lam"]), 250, mu=float(g["mu"]))
    assert rel(x_mid[0], st1.x0.ravel()) < TOL and rel(x_mid[2], st1.x2.ravel()) < TOL
    st = flat.spm_solve(g["s"], P, g["C"], np.ones(nb), g["g"], float(g["lam"]), 150, mu=float(g["mu"]), state=st1)
    assert rel(opt.x[0], st.x0.ravel()) < TOL and rel(opt.x[2], st.x2.ravel()) < TOL
    assert opt._mu[1, 0] == st.mu10 and opt._mu[2, 0] == st.mu20
    # warm start of a NEW optimizer from the exported state (pack of the folded state from canonical arrays)
    opt2 = O.SimpleOptimizer(model(), mu=float(g["mu"]))
    opt2.solve(250)
    opt3 = O.SimpleOptimizer(model(), mu=float(g["mu"]), x0=[np.array(v) for v in opt2.x])
    opt3._h[1, 0][:] = opt2._h[1, 0]
    opt3._h[2, 0][:] = opt2._h[2, 0]
    opt3._mu[:] = opt2._mu
    opt3.solve(150)
    assert rel(opt3.x[0], st.x0.ravel()) < 1e-9


def test_spm_packed_two_batch_axes(api):
    """rest_dims = (3, 2) (k-points x orbitals) is the same packed batch as rest_dims = (6,): fused engine, golden."""
    M, F, O = api
    g = golden("spm_packed")
    nb, L, Nw = 6, g["s"].size, g["P"].shape[0]
    rest = (3, 2)
    lstsq = F.ConstrainedLeastSquares(1.0, M.PartialDiagonalMatrix(-M.DiagonalMatrix(g["s"]), rest), g["g"].ravel(),
                                      M.PartialDiagonalMatrix(g["C"], rest), np.ones(nb))
    conds = [(0, 1, M.identity(L * nb), M.identity(L * nb)),
             (0, 2, M.PartialDiagonalMatrix(g["P"], rest), M.identity(Nw * nb))]
    opt = O.SimpleOptimizer(O.Model([lstsq, F.L1Regularizer(float(g["lam"]), L * nb), F.NonNegativePenalty(Nw * nb)], conds),
                            mu=float(g["mu"]))
    assert opt._plan_kind == "spm"
    opt.solve(400)
    assert rel(opt.x[0], g["x0"]) < TOL and rel(opt.x[1], g["x1"]) < TOL and rel(opt.x[2], g["x2"]) < TOL
    assert opt._mu[1, 0] == float(g["mu10"]) and opt._mu[2, 0] == float(g["mu20"])


def test_semi_positive_definite_penalty_golden(build_lib):
    """SemiPositiveDefinitePenalty.solve (objectivefunc.py:294-327) for every axis and
    mu type against the reference's outputs (tests/golden/psd.npz), plus the reference's own acceptance
    check (test_objectivefunc.py:169-185: all eigenvalues > -1e-10)."""
    from admmsolver_b200.matrix import DiagonalMatrix, PartialDiagonalMatrix, ScaledIdentityMatrix, identity
    from admmsolver_b200.objectivefunc import SemiPositiveDefinitePenalty
    g = golden("psd")
    N, K = 10, 20
    h = g["h"]
    p = SemiPositiveDefinitePenalty((N, N, K), axis=2)
    cases = [("x_identity", identity(N * N * K)),
             ("x_partial", PartialDiagonalMatrix(ScaledIdentityMatrix(N * N, 1.7), (K,))),
             ("x_partial_diag", PartialDiagonalMatrix(DiagonalMatrix(g["dvar"]), (K,))),
             ("x_diag", DiagonalMatrix(g["dfull"]))]
    for key, mu in cases:
        res = p.solve(h, mu)
        assert isinstance(res, np.ndarray) and res.dtype == np.float64
        assert rel(res, g[key]) < 1e-10, key
        x = res.reshape(N, N, K)
        for k in range(K):
            assert np.linalg.eigvalsh(x[:, :, k]).min() > -1e-10
    assert rel(SemiPositiveDefinitePenalty((6, 7, 7), axis=0).solve(g["h_axis0"], ScaledIdentityMatrix(6 * 7 * 7, 0.8)),
               g["x_axis0"]) < 1e-10
    assert rel(SemiPositiveDefinitePenalty((5, 4, 5), axis=1).solve(g["h_axis1"], ScaledIdentityMatrix(5 * 4 * 5, 1.3)),
               g["x_axis1"]) < 1e-10
    assert rel(SemiPositiveDefinitePenalty((3, 32, 32), axis=0).solve(g["h_n32"], ScaledIdentityMatrix(32 * 32 * 3, 1.0)),
               g["x_n32"]) < 1e-10


@pytest.mark.parametrize("n,K,axis", [(33, 5, 0), (48, 300, 2), (64, 3, 1), (100, 4, 0), (128, 2, 2), (160, 2, 0), (161, 2, 0),
                                      (200, 3, 2)])
def test_semi_positive_definite_penalty_large_slices(build_lib, n, K, axis):
    """SemiPositiveDefinitePenalty beyond 32 x 32 slices (the reference has no size limit: objectivefunc.py:294-327 loops
    np.linalg.eigh): the CTA-wide one-sided Jacobi on the shifted matrix -- shared memory up to n = 160, the L2-resident
    workspace beyond -- against the reference's own output (n = 40, golden) and the oracle restatement; random,
    already-PSD, negative definite, zero and +-lambda-paired (singular-value ties of the unshifted matrix) slices."""
    from admmsolver_b200.matrix import DiagonalMatrix, ScaledIdentityMatrix
    from admmsolver_b200.objectivefunc import SemiPositiveDefinitePenalty
    from oracle import flat
    g = golden("psd")
    if n == 33:
        res = SemiPositiveDefinitePenalty((40, 2, 40), axis=1).solve(g["h_n40"], ScaledIdentityMatrix(2 * 40 * 40, 0.6))
        assert rel(res, g["x_n40"]) < 1e-10
    rs = np.random.RandomState(n + K)
    shape = [n, n, n]
    shape[axis] = K
    X = rs.randn(K, n, n)
    X = X + np.swapaxes(X, 1, 2)
    B = rs.randn(n, n)
    X[0] = B @ B.T                                     # positive definite: unchanged by the projection
    if K > 1:
        X[1] = -(B @ B.T) - np.eye(n)                  # negative definite: projected to zero
    if K > 2:
        J = np.zeros((n, n))
        for i in range(0, n - 1, 2):
            J[i, i + 1] = J[i + 1, i] = 1.0 + i        # eigenvalue pairs +-(1 + i)
        X[2] = J
    if K > 3:
        X[3] = 0.0
    mu = np.linspace(0.5, 2.0, K * n * n)
    Xm = np.moveaxis(X, 0, axis)                       # (shape) with the slices along `axis`
    h = -(Xm.ravel() * mu) + 0.3j * rs.randn(K * n * n)           # -Re(h)/mu == X; the imaginary part is ignored
    res = SemiPositiveDefinitePenalty(tuple(shape), axis=axis).solve(h, DiagonalMatrix(mu))
    ref = flat.psd_project(h, mu, tuple(shape), axis)
    assert rel(res, ref) < 1e-10
    out = np.moveaxis(res.reshape(shape), axis, 0)
    assert rel(out[0], X[0]) < 1e-10
    if K > 1:
        assert np.abs(out[1]).max() < 1e-10 * np.abs(X[1]).max()
    if K > 3:
        assert np.all(out[3] == 0.0)
    for k in range(min(K, 6)):
        assert np.linalg.eigvalsh(out[k]).min() > -1e-10 * max(1.0, np.abs(X[k]).max())


def test_semi_positive_definite_model_golden(build_lib):
    """Matrix-valued least squares with a PSD constraint through SimpleOptimizer.solve (generic executor)."""
    from admmsolver_b200.matrix import identity
    from admmsolver_b200.objectivefunc import LeastSquares, SemiPositiveDefinitePenalty
    from admmsolver_b200.optimizer import Model, SimpleOptimizer
    g = golden("psd")
    n, k = 4, 3
    nx = n * n * k
    opt = SimpleOptimizer(Model([LeastSquares(0.9, g["loop_A"], g["loop_y"]), SemiPositiveDefinitePenalty((n, n, k), axis=2)],
                                [(0, 1, identity(nx), identity(nx))]), mu=0.5)
    opt.solve(120, interval_update_mu=25)
    assert rel(opt.x[0], g["loop_x0"]) < 1e-10 and rel(opt.x[1], g["loop_x1"]) < 1e-10
    assert float(opt._mu[1, 0]) == float(g["loop_mu10"])
    assert abs(opt(opt.x) - g["loop_objective"]) / abs(g["loop_objective"]) < 1e-10
    assert rel(opt._primal_residual, g["loop_primal"]) < 1e-8


def test_notebook_examples(build_lib):
    """examples/basis_pursuit.py and examples/spm.py: the two notebooks of the reference through the drop-in
    API.  Known answers: basis_pursuit.ipynb:137-138; spm.ipynb:270 (sum rule = 1)."""
    import importlib.util
    import os
    from conftest import ROOT

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    xanswer, x0, opt = load("basis_pursuit").main(verbose=False)
    assert abs(np.abs(xanswer).max() - 1.4312955709975443) < 1e-15
    assert abs(np.abs(xanswer - x0).max() - 0.0054070107628211295) < 1e-9
    for device_basis in (True, False):                              # IR basis by the on-device Jacobi SVD / by np.linalg.svd
        r = load("spm").main(niter=3000, verbose=False, device_basis=device_basis)
        assert r["L"] == 39 and abs(r["sum_rule"] - 1.0) < 1e-10       # spm.ipynb:214,270
        assert np.abs(r["rho_rec"] - r["rho"]).max() < 0.15 * r["rho"].max()      # the spectrum is recovered (noise 1e-4: the sharp peak is smoothed)
        assert r["rho_rec"].min() > -1e-3                               # and non-negative up to the ADMM residual


# ------------------------------------------------------------------ `_x_old` hand-over after fused solves
def _check_after(opt, g, tag, loose):
    """residual() / check_convergence() / update_mu() right after solve() equal the reference's (optimizer.py:232-299,324)."""
    assert len(opt._primal_residual) == len(g[f"{tag}_primal"])
    assert rel(opt.x[0], g[f"{tag}_x0"]) < TOL
    assert rel(opt._x_old[0], g[f"{tag}_x_old0"]) < 1e-9
    p, d = opt.residual()
    rp, rd = g[f"{tag}_residual_after"]
    # (absolute floor: the residuals of a converged run are rounding noise of the iterate, ~1e-13 here)
    assert abs(p - rp) <= 1e-7 * rp + 1e-14 and abs(d - rd) <= 1e-6 * rd + 1e-14, (tag, p, rp, d, rd)
    assert opt.check_convergence(1e-12) == bool(g[f"{tag}_converged_tight"])
    assert opt.check_convergence(loose) == bool(g[f"{tag}_converged_loose"])
    opt.update_mu()
    pairs = [(1, 0), (2, 0)][:len(g[f"{tag}_mu_after_update"])]
    assert [opt._mu[p_] for p_ in pairs] == list(g[f"{tag}_mu_after_update"])


def test_x_old_handover_after_fused_bp_solve(api):
    """VERDICT r01 / ADVICE: the step-wise public methods after a FUSED solve() (pattern A engine: cluster-resident
    kernel, streaming kernel, early exit, last iteration = mu-update iteration) against the reference's values."""
    M, F, O = api
    from admmsolver_b200 import problems
    g = golden("after_solve")

    def mk(A, y, lam):
        N = A.shape[1]
        o = O.SimpleOptimizer(O.Model([F.LeastSquares(1.0, A, y), F.L1Regularizer(lam, N)], [(1, 0, M.identity(N), M.identity(N))]))
        assert o._plan_kind == "bp"
        return o

    A, y, _ = problems.basis_pursuit_instance(100, 1000, 20, 1234)
    opt = mk(A, y, 0.1)
    opt.solve(100)
    assert opt._plan is not None                       # the fused engine ran, not the generic executor
    _check_after(opt, g, "bpnb", 1e-2)
    A, y, _ = problems.basis_pursuit_instance(128, 512, 10, 2)
    opt = mk(A, y, 0.1)
    opt.solve(301)
    _check_after(opt, g, "bp301", 1e-3)
    opt = mk(np.array([[2.0, 1.0]]), np.array([2.0]), 0.1)
    opt.solve(100)
    _check_after(opt, g, "lasso", 1e-6)


@pytest.mark.parametrize("tag,niter,nb", [("spm1", 250, None), ("spm6", 130, 6)])
def test_x_old_handover_after_fused_spm_solve(api, tag, niter, nb):
    """Same for the pattern B engine: single problem (cluster-resident kernel) and packed batch (batch kernels or
    co-resident clusters); the reference returns real values where round 1 raised AttributeError."""
    M, F, O = api
    g = golden("after_solve")
    s, P, Cm, gg = g[f"{tag}_s"], g[f"{tag}_P"], g[f"{tag}_C"], g[f"{tag}_g"]
    L, Nw = s.size, P.shape[0]
    lam, mu = float(g[f"{tag}_lam"]), float(g[f"{tag}_mu"])
    if nb is None:
        lstsq = F.ConstrainedLeastSquares(1.0, -M.DiagonalMatrix(s), gg, Cm, g[f"{tag}_D"])
        terms = [lstsq, F.L1Regularizer(lam, L), F.NonNegativePenalty(Nw)]
        conds = [(0, 1, M.identity(L), M.identity(L)), (0, 2, P, M.identity(Nw))]
    else:
        rest = (nb,)
        lstsq = F.ConstrainedLeastSquares(1.0, M.PartialDiagonalMatrix(-M.DiagonalMatrix(s), rest), gg.ravel(),
                                          M.PartialDiagonalMatrix(Cm, rest), g[f"{tag}_D"].astype(float))
        terms = [lstsq, F.L1Regularizer(lam, L * nb), F.NonNegativePenalty(Nw * nb)]
        conds = [(0, 1, M.identity(L * nb), M.identity(L * nb)), (0, 2, M.PartialDiagonalMatrix(P, rest), M.identity(Nw * nb))]
    opt = O.SimpleOptimizer(O.Model(terms, conds), mu=mu)
    assert opt._plan_kind == "spm"
    opt.solve(niter)
    assert opt._plan is not None
    _check_after(opt, g, tag, 1e-2)


def test_spm_x_old_from_batch_kernels(build_lib):
    """x0_old out of the batch kernels (x-update kernel / fused step kernel) equals the cluster-resident kernel's."""
    from admmsolver_b200 import problems
    from admmsolver_b200.batch import SharedSpM
    basis = problems.ir_basis()
    p = problems.spm_batch(5, basis, Nw=160, seed=2)
    outs = []
    for kw in (dict(), dict(mt=2, nsplit=1)):
        for solo in (False, True):
            if solo and kw:
                continue
            e = SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True, keep_x_old=True, **kw)
            e.solve(37, use_solo=solo)
            outs.append((e.x0_old(), e.x0()))
    for xo, x in outs[1:]:
        assert rel(xo, outs[0][0]) < 1e-11 and rel(x, outs[0][1]) < 1e-11
    assert rel(outs[0][0], outs[0][1]) > 1e-9              # really the previous iterate


def test_model_without_equality_conditions(api):
    """a model with no couplings does one sweep and stops (check_convergence() is vacuously True)."""
    M, F, O = api
    rs = np.random.RandomState(3)
    A, y = rs.randn(9, 4), rs.randn(9)
    opt = O.SimpleOptimizer(O.Model([F.LeastSquares(1.0, A, y)], []))
    opt.solve(50)
    assert opt._primal_residual == [0.0] and opt._dual_residual == [0.0]
    np.testing.assert_allclose(opt.x[0], np.linalg.lstsq(A, y, rcond=None)[0], atol=1e-12)


# ------------------------------------------------------------------ complex operators
def test_complex_operators_through_the_loop(api):
    """Complex A / C / D / coupling matrices (objectivefunc.py:76-77,89-96,138-157; matrix.py:100-118) against the
    reference's outputs: (a) complex LASSO; (b) a packed batch of 48 constrained problems sharing complex A, a two-row
    complex C and a dense complex coupling to a non-negative block.  The products run on the complex tensor-core GEMM,
    the cached factor (alpha A^H A + mu)^-1 on the Hermitian tensor-core inverse."""
    M, F, O = api
    g = golden("complex_ops")
    N = g["a_A"].shape[1]
    opt = O.SimpleOptimizer(O.Model([F.LeastSquares(0.8, g["a_A"], g["a_y"]), F.L1Regularizer(0.15, N)],
                                    [(1, 0, M.identity(N), M.identity(N))]))
    opt.solve(200, interval_update_mu=25)
    assert rel(opt.x[0], g["a_x0"]) < TOL and rel(opt.x[1], g["a_x1"]) < TOL
    assert opt._mu[1, 0] == g["a_mu10"]
    assert len(opt._primal_residual) == len(g["a_primal"]) and rel(opt._primal_residual, g["a_primal"]) < 1e-8
    assert abs(opt(opt.x) - g["a_objective"]) / g["a_objective"] < TOL

    nb = 48
    Lb, Nwb = g["b_A"].shape[1], g["b_P"].shape[0]
    rest = (nb,)
    terms = [F.ConstrainedLeastSquares(1.1, M.PartialDiagonalMatrix(g["b_A"], rest), g["b_y"],
                                       M.PartialDiagonalMatrix(g["b_C"], rest), g["b_D"]),
             F.L1Regularizer(0.3, Lb * nb), F.NonNegativePenalty(Nwb * nb)]
    conds = [(0, 1, M.identity(Lb * nb), M.identity(Lb * nb)),
             (0, 2, M.PartialDiagonalMatrix(g["b_P"], rest), M.identity(Nwb * nb))]
    opt = O.SimpleOptimizer(O.Model(terms, conds), mu=0.5)
    assert opt._plan_kind is None                     # complex operators: the generic executor
    opt.solve(160, interval_update_mu=20)
    for k in range(3):
        assert rel(opt.x[k], g[f"b_x{k}"]) < TOL, k
    assert rel(opt._h[2, 0], g["b_h20"]) < 1e-8
    assert opt._mu[1, 0] == g["b_mu10"] and opt._mu[2, 0] == g["b_mu20"]
    assert len(opt._primal_residual) == len(g["b_primal"]) and rel(opt._primal_residual, g["b_primal"]) < 1e-8
    assert abs(opt(opt.x) - g["b_objective"]) / abs(g["b_objective"]) < TOL
    # the constraint C x0 = D holds for every problem of the batch
    x0 = opt.x[0].reshape(Lb, nb)
    assert np.abs(g["b_C"] @ x0 - g["b_D"].reshape(2, nb)).max() < 1e-10


# ------------------------------------------------------------------ seeded model fuzz of the generic executor
@pytest.mark.parametrize("seed", range(14))
def test_generic_executor_model_fuzz_vs_reference(api, seed):
    """Random models (tests/golden/fuzz_models.py: 2-4 terms of every solvable kind, real / complex, coupled through
    identity, scaled identity, diagonal, dense rectangular and PartialDiagonalMatrix operators) through
    SimpleOptimizer.solve against the outputs of the unmodified reference on the same draws (fuzz_models.npz):
    every x block and the objective to 1e-9, identical penalties and iteration count (one case stops early),
    residual histories."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import fuzz_models
    g = golden("fuzz_models")
    opt, nterms = fuzz_models.build(api, seed)
    opt.solve(fuzz_models.NITER, interval_update_mu=fuzz_models.INTERVAL)
    assert len(opt._primal_residual) == len(g[f"s{seed}_primal"])
    for k in range(nterms):
        assert rel(opt.x[k], g[f"s{seed}_x{k}"]) < 1e-9, (seed, k)
    assert [opt._mu[k, 0] for k in range(1, nterms)] == list(g[f"s{seed}_mu"])
    assert rel(opt._primal_residual, g[f"s{seed}_primal"]) < 1e-8 and rel(opt._dual_residual, g[f"s{seed}_dual"]) < 1e-8
    ref_obj = float(g[f"s{seed}_objective"])
    assert abs(opt(opt.x) - ref_obj) <= 1e-9 * abs(ref_obj)


def test_spm_several_constraint_rows_dropin(api):
    """The drop-in API with a constraint MATRIX (three rows; packed: two rows): recognised as the SpM pattern and run
    on the fused engine (not the generic executor), equal to the reference (spm_multirow.npz)."""
    M, F, O = api
    g = golden("spm_multirow")
    L, Nw = g["a_s"].size, g["a_P"].shape[0]
    lam, mu = float(g["a_lam"]), float(g["a_mu"])
    lstsq = F.ConstrainedLeastSquares(1.0, -M.DiagonalMatrix(g["a_s"]), g["a_g"], g["a_C"], g["a_D"])
    opt = O.SimpleOptimizer(O.Model([lstsq, F.L1Regularizer(lam, L), F.NonNegativePenalty(Nw)],
                                    [(0, 1, M.identity(L), M.identity(L)), (0, 2, g["a_P"], M.identity(Nw))]), mu=mu)
    assert opt._plan_kind == "spm"
    opt.solve(300, interval_update_mu=50)
    for k in range(3):
        assert rel(opt.x[k], g[f"a_x{k}"]) < TOL
    assert opt._mu[1, 0] == float(g["a_mu10"]) and opt._mu[2, 0] == float(g["a_mu20"])
    assert abs(opt(opt.x) - float(g["a_objective"])) <= TOL * abs(float(g["a_objective"]))
    nb = 6
    rest = (nb,)
    lstsq = F.ConstrainedLeastSquares(1.0, M.PartialDiagonalMatrix(-M.DiagonalMatrix(g["a_s"]), rest), g["b_g"].ravel(),
                                      M.PartialDiagonalMatrix(g["b_C"], rest), g["b_D"].ravel())
    opt = O.SimpleOptimizer(O.Model([lstsq, F.L1Regularizer(lam, L * nb), F.NonNegativePenalty(Nw * nb)],
                                    [(0, 1, M.identity(L * nb), M.identity(L * nb)),
                                     (0, 2, M.PartialDiagonalMatrix(g["a_P"], rest), M.identity(Nw * nb))]), mu=mu)
    assert opt._plan_kind == "spm"
    opt.solve(250, interval_update_mu=50)
    for k in range(3):
        assert rel(opt.x[k], g[f"b_x{k}"]) < TOL
    assert opt._mu[1, 0] == float(g["b_mu10"]) and opt._mu[2, 0] == float(g["b_mu20"])
    assert len(opt._primal_residual) == 250 and rel(opt._primal_residual, g["b_primal"]) < 1e-8


@pytest.mark.parametrize("moments", [False, True])
def test_spm_batch_pipeline_example(build_lib, moments):
    """examples/spm_batch.py: the whole analytic-continuation batch on the device -- basis by Jacobi SVD, G(tau) -> g_l,
    fused engine with per-problem criterion (one or two constraint rows), reconstruction -- recovers the model spectra,
    keeps them non-negative and satisfies the constraints to rounding."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("spm_batch", os.path.join(ROOT, "examples", "spm_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.main(nb=40, niter=1500, nw=400, moments=moments, verbose=False)
    assert r["L"] == 39 and r["folded"]                             # (device basis, P projected onto exact parity)
    assert r["constraint_violation"] < 1e-9
    assert r["min_rho"] > -1e-2
    assert r["data_misfit"] < 3e-2                                  # the Green's functions are reproduced (L1 weight 1e-5, 1500 iterations)
    assert r["median_rel_err"] < 0.3 and r["max_rel_err"] < 0.9     # (analytic continuation is ill-posed: sharp peaks are smoothed)


def test_semi_positive_definite_model_large_slice_golden(build_lib):
    """Matrix-valued least squares with a PSD constraint on ONE 36 x 36 slice (1296 unknowns) through
    SimpleOptimizer.solve: the generic executor with the CTA-wide Jacobi projection inside the loop (graph-captured
    iterations), the cached factor of the 1296 x 1296 least-squares term -- against the reference (psd.npz, loop36_*)."""
    from admmsolver_b200.matrix import identity
    from admmsolver_b200.objectivefunc import LeastSquares, SemiPositiveDefinitePenalty
    from admmsolver_b200.optimizer import Model, SimpleOptimizer
    g = golden("psd")
    rs36 = np.random.RandomState(36)
    n36 = 36
    nx = n36 * n36
    A36 = rs36.randn(nx + 200, nx) / np.sqrt(nx)
    y36 = rs36.randn(nx + 200)
    opt = SimpleOptimizer(Model([LeastSquares(1.0, A36, y36), SemiPositiveDefinitePenalty((n36, n36, 1), axis=2)],
                                [(0, 1, identity(nx), identity(nx))]), mu=0.5)
    opt.solve(60, interval_update_mu=20)
    assert rel(opt.x[0], g["loop36_x0"]) < 1e-9 and rel(opt.x[1], g["loop36_x1"]) < 1e-9
    assert float(opt._mu[1, 0]) == float(g["loop36_mu10"])
    assert rel(opt._primal_residual, g["loop36_primal"]) < 1e-8
    assert abs(opt(opt.x) - g["loop36_objective"]) / abs(g["loop36_objective"]) < 1e-9
    assert np.linalg.eigvalsh(opt.x[1].real.reshape(n36, n36)).min() > -1e-10
