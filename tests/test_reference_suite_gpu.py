"""GPU: the reference's own acceptance checks, one for one, run against the drop-in package.

Every test states the independent ground truth the corresponding reference test uses (a generic
scipy minimiser, a closed form, a defining property) -- no golden files, no oracle: a user who
swaps ``admmsolver`` for ``admmsolver_b200`` and runs the reference's suite sees exactly these checks.
Reference: /root/reference/test/test_{matrix,objectivefunc,optimizer,util}.py (file:line per test)."""
import numpy as np
import pytest
import torch
from scipy.optimize import minimize

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(build_lib):
    assert torch.cuda.is_available()
    import admmsolver_b200.matrix as M
    import admmsolver_b200.objectivefunc as F
    import admmsolver_b200.optimizer as O
    import admmsolver_b200.util as U
    return M, F, O, U


def _crandn(rs, *shape):
    return rs.randn(*shape) + 1j * rs.randn(*shape)


def _argmin_complex(f, x0):
    """BFGS over (Re, Im) of a real-valued function of a complex vector."""
    pack = lambda z: np.concatenate([z.real, z.imag])
    unpack = lambda v: v[:v.size // 2] + 1j * v[v.size // 2:]
    res = minimize(lambda v: f(unpack(v)), pack(np.asarray(x0, dtype=complex)), method="BFGS",
                   options={"maxiter": 100000})
    return unpack(res.x)


def _quadratic(h, mu):
    """h^H x + x^H h + x^H mu x, the coupling part every term's solve() minimises (objectivefunc.py:41-53)."""
    md = mu.asmatrix()
    return lambda x: float(np.real(2 * np.vdot(h, x) + np.vdot(x, md @ x)))


# ------------------------------------------------------------------ test_matrix.py
def _zoo(M, rs, n, rest):
    """One matrix of every type, all n x n with n = 3 * rest."""
    return [M.DiagonalMatrix(_crandn(rs, n)), M.ScaledIdentityMatrix(n, 1.5 - 0.5j),
            M.PartialDiagonalMatrix(_crandn(rs, 3, 3), rest_dims=(rest,)), M.DenseMatrix(_crandn(rs, n, n))]


def test_matmul(pkg):
    """test_matrix.py:11-33 (left square, right n x k of every type) and :36-57 (left k x n)."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n, rest = 12, 4
    tall = [M.DenseMatrix(_crandn(rs, n, 4)), M.ScaledIdentityMatrix((n, 4), 1 + 1j),
            M.PartialDiagonalMatrix(_crandn(rs, 3, 1), rest_dims=(rest,))]
    wide = [M.DenseMatrix(_crandn(rs, 4, n)), M.PartialDiagonalMatrix(_crandn(rs, 1, 3), rest_dims=(rest,))]
    for sq in _zoo(M, rs, n, rest):
        for t in tall:
            p = sq @ t
            assert isinstance(p, M.MatrixBase) and p.shape == (n, 4)
            np.testing.assert_allclose(p.asmatrix(), sq.asmatrix() @ t.asmatrix(), atol=1e-12)
        for w in wide:
            p = w @ sq
            assert isinstance(p, M.MatrixBase) and p.shape == (4, n)
            np.testing.assert_allclose(p.asmatrix(), w.asmatrix() @ sq.asmatrix(), atol=1e-12)


def test_mul_transpose_conj(pkg):
    """test_matrix.py:60-88: scalar *, .T, .conjugate() agree with the dense matrix for every type."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    for m in _zoo(M, rs, 4 * 3, 4) + [M.DiagonalMatrix(_crandn(rs, 2), shape=(4, 2))]:
        d = m.asmatrix()
        for c in (2.0, 1j, 0.5 - 2j):
            np.testing.assert_allclose((m * c).asmatrix(), d * c, atol=1e-14)
            np.testing.assert_allclose((c * m).asmatrix(), c * d, atol=1e-14)
        np.testing.assert_allclose(m.T.asmatrix(), d.T, atol=0)
        np.testing.assert_allclose(m.conjugate().asmatrix(), d.conjugate(), atol=0)
        np.testing.assert_allclose(m.conj().T.asmatrix(), d.conj().T, atol=0)


def test_add(pkg):
    """test_matrix.py:91-107: + of every pair of types."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    zoo = _zoo(M, rs, 12, 4)
    for a in zoo:
        for b in zoo:
            np.testing.assert_allclose((a + b).asmatrix(), a.asmatrix() + b.asmatrix(), atol=1e-13)


def test_DiagonalMatrix_PartialDiagonalMatrix(pkg):
    """test_matrix.py:110-124: a diagonal that is constant along the rest axes + PartialDiagonal (two rest axes)
    keeps the Kronecker type."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n = 3
    dg = M.DiagonalMatrix(np.repeat(rs.randn(n), 4))
    pd = M.PartialDiagonalMatrix(_crandn(rs, n, n), (2, 2))
    s = dg + pd
    assert isinstance(s, M.PartialDiagonalMatrix)
    np.testing.assert_allclose(s.asmatrix(), dg.asmatrix() + pd.asmatrix(), atol=1e-14)
    np.testing.assert_allclose((pd + dg).asmatrix(), dg.asmatrix() + pd.asmatrix(), atol=1e-14)


def test_PartialDiagonalMatrix_PartialDiagonalMatrix(pkg):
    """test_matrix.py:127-134."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    a, b = M.PartialDiagonalMatrix(_crandn(rs, 3, 3), (2, 2)), M.PartialDiagonalMatrix(_crandn(rs, 3, 3), (2, 2))
    s = a + b
    assert isinstance(s, M.PartialDiagonalMatrix)
    np.testing.assert_allclose(s.asmatrix(), a.asmatrix() + b.asmatrix(), atol=1e-14)


def test_matmul_DiagonalMatrix_PartialDiagonalMatrix(pkg):
    """test_matrix.py:137-150: Diagonal @ PartialDiagonal stays PartialDiagonal."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n = 3
    dg = M.DiagonalMatrix(np.repeat(rs.randn(n), 4))
    pd = M.PartialDiagonalMatrix(_crandn(rs, n, n), (2, 2))
    p = dg @ pd
    assert isinstance(p, M.PartialDiagonalMatrix)
    np.testing.assert_allclose(p.asmatrix(), dg.asmatrix() @ pd.asmatrix(), atol=1e-14)
    np.testing.assert_allclose((pd @ dg).asmatrix(), pd.asmatrix() @ dg.asmatrix(), atol=1e-14)


def test_inv(pkg):
    """test_matrix.py:153-166: inv() of every square type is a MatrixBase and a two-sided inverse."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n = 12
    for m in _zoo(M, rs, n, 4):
        mi = m.inv()
        assert isinstance(mi, M.MatrixBase)
        np.testing.assert_allclose(mi.asmatrix() @ m.asmatrix(), np.identity(n), atol=1e-10)
        np.testing.assert_allclose(m.asmatrix() @ mi.asmatrix(), np.identity(n), atol=1e-10)


@pytest.mark.parametrize("shape", [(4, 4), (2, 4), (4, 2)])
def test_matvec(pkg, shape):
    """test_matrix.py:169-210: M @ v for a 1-D v, square and rectangular operators of every type."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n, m = shape
    k = min(n, m)
    ops = [M.DiagonalMatrix(_crandn(rs, k), shape=shape), M.ScaledIdentityMatrix(shape, 1 + 1j),
           M.PartialDiagonalMatrix(_crandn(rs, n // 2, m // 2), (2,)), M.DenseMatrix(_crandn(rs, n, m))]
    v = _crandn(rs, m)
    for op in ops:
        out = op @ v
        assert isinstance(out, np.ndarray) and out.shape == (n,)
        np.testing.assert_allclose(out, op.asmatrix() @ v, atol=1e-13)


def test_batched_matvec(pkg):
    """test_matrix.py:213-233: M @ V with V of shape (m, nbatch)."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    n, nbatch = 4, 3
    V = _crandn(rs, n, nbatch)
    for op in [M.DiagonalMatrix(_crandn(rs, n)), M.ScaledIdentityMatrix(n, 2.0), M.PartialDiagonalMatrix(_crandn(rs, 2, 2), (2,)),
               M.DenseMatrix(_crandn(rs, n, n))]:
        np.testing.assert_allclose(op @ V, op.asmatrix() @ V, atol=1e-13)


def test_matmul_diagonal(pkg):
    """test_matrix.py:236-244: (4 x 2 diagonal) @ (2 x 4 diagonal) is the zero-padded elementwise product."""
    M = pkg[0]
    rs = np.random.RandomState(100)
    a, b = M.DiagonalMatrix(rs.randn(2), shape=(4, 2)), M.DiagonalMatrix(rs.randn(2), shape=(2, 4))
    ab = a @ b
    assert isinstance(ab, M.DiagonalMatrix) and ab.shape == (4, 4)
    np.testing.assert_allclose(ab.diagonals, np.concatenate([a.diagonals * b.diagonals, np.zeros(2)]))


def test_vecprod_and_pad_by_zero(pkg):
    """test_matrix.py:247-257."""
    M = pkg[0]
    np.testing.assert_allclose(M._vecprod(np.ones(1), np.ones(2), 3), [1, 0, 0])
    np.testing.assert_allclose(M._pad_by_zero(np.ones(1), 3), [1, 0, 0])


# ------------------------------------------------------------------ test_objectivefunc.py
@pytest.mark.parametrize("partial", [False, True])
def test_least_squares(pkg, partial):
    """test_objectivefunc.py:34-53 (dense A) and :56-79 (A = PartialDiagonalMatrix): solve(h, mu) with a dense
    Hermitian mu is the minimiser of alpha |y - A x|^2 + h^H x + x^H h + x^H mu x found by BFGS."""
    M, F = pkg[0], pkg[1]
    rs = np.random.RandomState(100)
    if partial:
        n1, n2, rest = 40, 20, 20
        A = M.PartialDiagonalMatrix(_crandn(rs, n1 // rest, n2 // rest), rest_dims=(rest,))
        assert A.shape == (n1, n2)
    else:
        n1, n2 = 4, 2
        A = _crandn(rs, n1, n2)
    alpha = 2.0
    y, h = _crandn(rs, n1), _crandn(rs, n2)
    r = _crandn(rs, n2, n2)
    mu = M.asmatrixtype(r.conj().T @ r)
    x = F.LeastSquares(alpha, A, y).solve(h, mu)
    Ad = A.asmatrix() if partial else A
    quad = _quadratic(h, mu)
    f = lambda z: alpha * np.linalg.norm(y - Ad @ z) ** 2 + quad(z)
    x_ref = _argmin_complex(f, x)
    np.testing.assert_allclose(x, x_ref, rtol=1e-4 if partial else 1e-8, atol=1e-7)
    np.testing.assert_allclose(f(x), f(x_ref), rtol=1e-8)
    assert f(x) <= f(x_ref) + 1e-9 * abs(f(x_ref))


def test_constrained_least_squares(pkg):
    """test_objectivefunc.py:82-103: C x = D holds to 1e-10; in addition (the reference leaves this as a FIXME)
    x minimises the objective on the constraint set: compared with BFGS over the null space of C."""
    M, F = pkg[0], pkg[1]
    rs = np.random.RandomState(100)
    n1, n2, nc = 8, 4, 2
    alpha = 2.0
    y, A, h = _crandn(rs, n1), _crandn(rs, n1, n2), _crandn(rs, n2)
    C_, D = _crandn(rs, nc, n2), _crandn(rs, nc)
    r = _crandn(rs, n2, n2)
    mu = M.asmatrixtype(r.conj().T @ r)
    x = F.ConstrainedLeastSquares(alpha, A, y, C_, D).solve(h, mu)
    assert np.abs(C_ @ x - D).max() < 1e-10
    xp = np.linalg.lstsq(C_, D, rcond=None)[0]
    null = np.linalg.svd(C_)[2][nc:].conj().T                 # n2 x (n2 - nc)
    quad = _quadratic(h, mu)
    f = lambda z: alpha * np.linalg.norm(y - A @ z) ** 2 + quad(z)
    t = _argmin_complex(lambda t: f(xp + null @ t), null.conj().T @ (x - xp))
    np.testing.assert_allclose(x, xp + null @ t, rtol=1e-6, atol=1e-7)


def test_L1(pkg):
    """test_objectivefunc.py:106-123: elementwise minimiser of alpha |x| + 2 h x + mu x^2."""
    M, F = pkg[0], pkg[1]
    n = 20
    h = 0.5 * np.arange(-n // 2, n // 2)
    mu = M.identity(n)
    alpha = 1.0
    x = F.L1Regularizer(alpha, n).solve(h, mu)
    for i in range(n):
        res = minimize(lambda t: alpha * np.abs(t) + 2 * h[i] * t + mu.diagonals[i] * t ** 2, 0.0, method="BFGS")
        assert abs(x[i] - res.x[0]) < 1e-5


def test_non_negative(pkg):
    """test_objectivefunc.py:126-144: minimiser of 2 h x + mu x^2 on x >= 0 (closed form max(0, -h/mu) and the
    reference's penalised BFGS)."""
    M, F = pkg[0], pkg[1]
    h = np.array([0.0, -10.0, 10.0])
    mu = M.identity(h.size)
    x = F.NonNegativePenalty(h.size).solve(h, mu)
    np.testing.assert_allclose(x, np.maximum(0.0, -h / mu.diagonals), atol=1e-14)
    for i in range(h.size):
        res = minimize(lambda t: 1e5 * max(-t, 0.0) + 2 * h[i] * t + mu.diagonals[i] * t ** 2, 0.0, method="BFGS")
        assert abs(x[i] - res.x[0]) < 1e-5


def test_L2(pkg):
    """test_objectivefunc.py:147-166: minimiser of alpha |A x|^2 + 2 Re h^H x + x^H mu x (BFGS and closed form)."""
    M, F = pkg[0], pkg[1]
    rs = np.random.RandomState(7)
    n, m = 10, 5
    r = _crandn(rs, n, n)
    mu = M.asmatrixtype(r.conj().T @ r)
    alpha = 2.0
    A, h = _crandn(rs, m, n), _crandn(rs, n)
    x = F.L2Regularizer(alpha, A).solve(h, mu)
    np.testing.assert_allclose(x, -np.linalg.solve(alpha * A.conj().T @ A + mu.asmatrix(), h), rtol=1e-9, atol=1e-12)
    quad = _quadratic(h, mu)
    f = lambda z: alpha * np.linalg.norm(A @ z) ** 2 + quad(z)
    x_ref = _argmin_complex(f, x)
    np.testing.assert_allclose(x, x_ref, atol=np.abs(x_ref).max() * 1e-5, rtol=0)


def test_semi_positive_definite_penalty(pkg):
    """test_objectivefunc.py:169-185: every N x N slice along the batch axis comes back positive semi-definite,
    for mu = identity and mu = PartialDiagonal(ScaledIdentity)."""
    M, F = pkg[0], pkg[1]
    rs = np.random.RandomState(100)
    K, N = 20, 10
    h = _crandn(rs, N * N * K)
    for mu in (M.asmatrixtype(M.identity(N * N * K)), M.PartialDiagonalMatrix(M.ScaledIdentityMatrix(N * N, 1.0), (K,))):
        x = np.asarray(F.SemiPositiveDefinitePenalty((N, N, K), axis=2).solve(h, mu)).reshape(N, N, K)
        for k in range(K):
            assert np.all(np.linalg.eigvalsh(x[:, :, k]) > -1e-10)


# ------------------------------------------------------------------ test_optimizer.py
def test_LASSO(pkg):
    """test_optimizer.py:13-50: 1 x 2 LASSO; ADMM reaches the Nelder-Mead minimiser to 1e-10 in 100 iterations and
    the objective of the model equals the plain function value."""
    M, F, O = pkg[0], pkg[1], pkg[2]
    y, A, alpha = np.array([2]), np.array([[2, 1]]), 0.1
    f = lambda x: np.linalg.norm(y - A @ x) ** 2 + alpha * np.sum(np.abs(x))
    res = minimize(f, x0=np.array([1.1, 0]), method="Nelder-Mead", options={"xatol": 1e-10})
    assert res.success
    model = O.Model([F.LeastSquares(1.0, A, y), F.L1Regularizer(alpha, 2)], [(1, 0, M.identity(2), M.identity(2))])
    opt = O.SimpleOptimizer(model)
    assert abs(opt(2 * [res.x]) - f(res.x)) < 1e-10
    opt.solve(100)
    for x in opt.x:
        np.testing.assert_allclose(x, res.x, atol=1e-10)


def test_basis_pursuit(pkg):
    """test_optimizer.py:52-82: 100 x 1000 Gaussian A, 20-sparse signal recovered to 1 % of its largest entry
    after 100 iterations."""
    M, F, O = pkg[0], pkg[1], pkg[2]
    N, Mrows, K = 1000, 100, 20
    np.random.seed(1234)
    A = np.random.randn(Mrows, N)
    xanswer = np.zeros(N)
    xanswer[:K] = np.random.randn(K)
    xanswer = np.random.permutation(xanswer)
    model = O.Model([F.LeastSquares(1.0, A, A @ xanswer), F.L1Regularizer(1e-1, N)], [(1, 0, M.identity(N), M.identity(N))])
    opt = O.SimpleOptimizer(model)
    opt.solve(100)
    np.testing.assert_allclose(opt.x[0], xanswer, atol=1e-2 * np.abs(xanswer).max(), rtol=0)


def test_ridge(pkg):
    """test_optimizer.py:85-109: complex ridge regression converges to (A^H A + alpha B^H B)^-1 A^H y."""
    M, F, O = pkg[0], pkg[1], pkg[2]
    rs = np.random.RandomState(100)
    y, A, B = _crandn(rs, 2), _crandn(rs, 2, 2), _crandn(rs, 1, 2)
    alpha = 1
    model = O.Model([F.LeastSquares(1.0, A, y), F.L2Regularizer(alpha, B)], [(1, 0, M.identity(2), M.identity(2))])
    opt = O.SimpleOptimizer(model)
    opt.solve(niter=100, update_h=True)
    x_ref = np.linalg.inv(A.conj().T @ A + alpha * B.conj().T @ B) @ A.conj().T @ y
    np.testing.assert_allclose(opt.x[0], x_ref, atol=np.abs(x_ref).max() * 1e-8)


# ------------------------------------------------------------------ test_util.py
def test_second_deriv_prj(pkg):
    """test_util.py:3-16: the three-point projector gives f'' = 2 for f = x^2 on a non-uniform mesh."""
    U = pkg[3]
    x = np.linspace(0, np.sqrt(3), 1000) ** 2
    np.testing.assert_allclose(U.second_deriv_prj(x) @ (x ** 2), np.full(998, 2.0))


def test_smooth_regularizer_coeff(pkg):
    """test_util.py:19-36: |prj @ w^2|^2 = integral of (f'')^2 = 4 (w_max - w_min)."""
    U = pkg[3]
    omega = np.linspace(0.0, np.sqrt(3.0), 10000) ** 2
    assert abs(np.linalg.norm(U.smooth_regularizer_coeff(omega) @ omega ** 2) ** 2 - 3.0 * 4.0) < 1e-2


# ------------------------------------------------------------------ error behaviour
def test_error_behaviour_matches_reference(pkg):
    """The exceptions a caller of the reference can rely on: RuntimeError for duplicate conditions
    (optimizer.py:111-112), for inverses / diagonals of rectangular matrices (matrix.py:157,167,225); ValueError for an
    invalid shape (matrix.py:142); AssertionError for a missing mu (objectivefunc.py:262-263), shape mismatches between
    conditions and terms (optimizer.py:107-110), non-float coefficients (matrix.py:134-135) and wrong argument types."""
    M, F, O, U = pkg
    lst, l1 = F.LeastSquares(1.0, np.ones((3, 4)), np.ones(3)), F.L1Regularizer(0.1, 4)
    with pytest.raises(RuntimeError):
        O.Model([lst, l1], [(1, 0, M.identity(4), M.identity(4)), (1, 0, M.identity(4), M.identity(4))])
    with pytest.raises(AssertionError):
        O.Model([lst, l1], [(1, 0, M.identity(5), M.identity(5))])
    with pytest.raises(RuntimeError):
        M.ScaledIdentityMatrix((2, 3), 1.0).inv()
    with pytest.raises(RuntimeError):
        M.ScaledIdentityMatrix((2, 3), 1.0).diagonals
    with pytest.raises(RuntimeError):
        M.DiagonalMatrix(np.ones(2), shape=(3, 2)).inv()
    with pytest.raises(ValueError):
        M.ScaledIdentityMatrix([2, 2], 1.0)
    with pytest.raises(AssertionError):
        M.ScaledIdentityMatrix(2, 1)
    with pytest.raises(AssertionError):            # the ValueError of objectivefunc.py:270 is unreachable: :263 asserts first
        F.NonNegativePenalty(3).solve(np.zeros(3), None)
    with pytest.raises(AssertionError):
        F.L1Regularizer(0.1, 3).solve(np.zeros(3), M.DenseMatrix(np.eye(3)))
    with pytest.raises(AssertionError):
        F.LeastSquares(1.0, np.ones((3, 4)), np.ones(5))
    # a condition is a 4-tuple or an EqualityCondition with .size (optimizer.py:12-38)
    ec = O.EqualityCondition(1, 0, M.identity(4), M.identity(4))
    assert (ec.i1, ec.i2, ec.size) == (1, 0, 4)
    m = O.Problem([lst, l1], [ec])
    assert m.num_func == 2 and m.E[1, 0] is not None and m.E[0, 1] is not None
