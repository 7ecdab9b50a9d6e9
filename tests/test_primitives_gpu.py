"""Device primitives behind the C ABI against NumPy/LAPACK (float64, tolerance in each test)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 5, 8, 37, 64, 100, 128, 150, 200, 256, 300, 512, 520])
def test_spd_inverse_batched(build_lib, n):
    """admm_spd_inverse_batched: tensor-core block Gauss-Jordan (register resident for n <= 128, L2 resident for n <= 512) and the scalar
    fallback (n > 512)
    vs np.linalg.inv on K = A A^T + mu I (the Woodbury matrix of basis pursuit), masked batch, info flags.
    Replaces np.linalg.inv at matrix.py:77-78 via objectivefunc.py:89-96."""
    from admmsolver_b200 import _lib
    rs = np.random.RandomState(n)
    nbatch = 7
    mats = []
    for b in range(nbatch):
        A = rs.randn(n, 3 * n + 2)
        mats.append(A @ A.T + (0.5 + b) * np.eye(n))
    K = np.stack(mats)
    lda = n + 3                                        # non-trivial leading dimension
    buf = np.zeros((nbatch, n, lda))
    buf[:, :, :n] = K
    d = torch.from_numpy(buf).cuda()
    mask = torch.tensor([1, 1, 0, 1, 1, 1, 1], dtype=torch.int32, device="cuda")
    info = torch.full((nbatch,), -1, dtype=torch.int32, device="cuda")
    _lib.call("admm_spd_inverse_batched", n, nbatch, _lib.ptr(d), n * lda, lda, _lib.ptr(mask), _lib.ptr(info), _lib.stream())
    out = d.cpu().numpy()
    inf = info.cpu().numpy()
    for b in range(nbatch):
        if b == 2:                                     # masked out: untouched
            assert np.array_equal(out[b], buf[b]) and inf[b] == -1
            continue
        ref = np.linalg.inv(K[b])
        err = np.linalg.norm(out[b, :, :n] - ref) / np.linalg.norm(ref)
        assert err < 1e-12, (n, b, err)
        assert np.array_equal(out[b, :, n:], buf[b, :, n:])          # padding columns untouched
        assert inf[b] == 0
        # K^-1 K = I to working precision (test_matrix.py:153 uses atol 1e-12 on inv @ m)
        assert np.abs(out[b, :, :n] @ K[b] - np.eye(n)).max() < 1e-10


def test_spd_inverse_flags_indefinite(build_lib):
    from admmsolver_b200 import _lib
    K = np.eye(16)
    K[5, 5] = -1.0
    d = torch.from_numpy(K.copy()).cuda()
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("admm_spd_inverse_batched", 16, 1, _lib.ptr(d), 256, 16, None, _lib.ptr(info), _lib.stream())
    assert int(info.item()) == 6                       # 1-based index of the first non-positive pivot


@pytest.mark.parametrize("m,n,k", [(64, 64, 16), (70, 100, 33), (1000, 4096, 1000), (39, 2000, 39), (33, 32, 16), (257, 65, 129)])
@pytest.mark.parametrize("op", [0, 1])
def test_gemm_tensor_core(build_lib, m, n, k, op):
    """admm_gemm, real FP64 on the tensor cores (DMMA) incl. ragged edges and the transposed-A form, vs NumPy.
    Replaces DenseMatrix.__matmul__ / the tensordot of _matvec_impl (matrix.py:100-118,392-397)."""
    from admmsolver_b200 import _lib
    rs = np.random.RandomState(m + n + k + op)
    A = rs.randn(m, k) if op == 0 else rs.randn(k, m)
    B = rs.randn(k, n)
    ref = (A if op == 0 else A.T) @ B
    lda, ldb, ldc = A.shape[1] + 3, n + 1, n + 5
    Ad = torch.zeros(A.shape[0], lda, dtype=torch.float64, device="cuda")
    Ad[:, :A.shape[1]] = torch.from_numpy(A).cuda()
    Bd = torch.zeros(k, ldb, dtype=torch.float64, device="cuda")
    Bd[:, :n] = torch.from_numpy(B).cuda()
    Cd = torch.full((m, ldc), 7.0, dtype=torch.float64, device="cuda")
    _lib.call("admm_gemm", 0, op, m, n, k, _lib.ptr(Ad), lda, _lib.ptr(Bd), ldb, _lib.ptr(Cd), ldc, _lib.stream())
    out = Cd.cpu().numpy()
    assert np.abs(out[:, :n] - ref).max() <= 1e-12 * np.abs(A).max() * np.abs(B).max() * k
    assert np.all(out[:, n:] == 7.0)                  # nothing written outside the m x n block


@pytest.mark.parametrize("m,n,k", [(64, 64, 16), (70, 100, 33), (300, 1024, 300), (39, 500, 39), (33, 16, 8), (257, 65, 129)])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_gemm_complex_tensor_core(build_lib, m, n, k, op):
    """admm_gemm on complex128 operands: the real form of the product on the tensor cores (gemm_dmma_kernel<OP, true>,
    no operand is split or copied), plain / transposed / conjugate-transposed A, ragged edges and padded leading
    dimensions, vs NumPy.  Replaces DenseMatrix.__matmul__ with complex A (matrix.py:100-118) and the A^H A / A^H y
    products of a complex LeastSquares term (objectivefunc.py:76-77,101-110)."""
    from admmsolver_b200 import _lib
    rs = np.random.RandomState(m + n + k + op)
    shp = (m, k) if op == 0 else (k, m)
    A = rs.randn(*shp) + 1j * rs.randn(*shp)
    B = rs.randn(k, n) + 1j * rs.randn(k, n)
    ref = (A if op == 0 else (A.T if op == 1 else A.conj().T)) @ B
    lda, ldb, ldc = shp[1] + 3, n + 1, n + 5
    Ad = torch.zeros(shp[0], lda, dtype=torch.complex128, device="cuda")
    Ad[:, :shp[1]] = torch.from_numpy(A).cuda()
    Bd = torch.zeros(k, ldb, dtype=torch.complex128, device="cuda")
    Bd[:, :n] = torch.from_numpy(B).cuda()
    Cd = torch.full((m, ldc), 7.0 - 3.0j, dtype=torch.complex128, device="cuda")
    _lib.call("admm_gemm", 1, op, m, n, k, _lib.ptr(Ad), lda, _lib.ptr(Bd), ldb, _lib.ptr(Cd), ldc, _lib.stream())
    out = Cd.cpu().numpy()
    assert np.abs(out[:, :n] - ref).max() <= 4e-12 * np.abs(A).max() * np.abs(B).max() * k
    assert np.all(out[:, n:] == 7.0 - 3.0j)           # nothing written outside the m x n block


@pytest.mark.parametrize("n", [1, 3, 8, 20, 39, 64, 100, 200, 256, 300])
def test_hpd_inverse_batched(build_lib, n):
    """admm_hpd_inverse_batched: complex Hermitian positive definite inverse through the real form of order 2n on the
    tensor-core SPD kernels (register resident for n <= 64, L2 resident for n <= 256, scalar beyond) vs np.linalg.inv on
    G = A^H A + mu I with complex A; masked batch, padded leading dimension, info flags, indefinite input untouched."""
    from admmsolver_b200 import _lib
    rs = np.random.RandomState(100 + n)
    nbatch = 5
    mats = []
    for b in range(nbatch):
        A = rs.randn(2 * n + 3, n) + 1j * rs.randn(2 * n + 3, n)
        mats.append(A.conj().T @ A + (0.5 + b) * np.eye(n))
    mats[3] = mats[3] - 1e3 * np.eye(n) * (1 + np.abs(mats[3]).max())          # indefinite
    G = np.stack(mats)
    lda = n + 2
    buf = np.zeros((nbatch, n, lda), dtype=np.complex128)
    buf[:, :, :n] = G
    d = torch.from_numpy(buf).cuda()
    work = torch.empty(nbatch * 4 * n * n, dtype=torch.float64, device="cuda")
    mask = torch.tensor([1, 0, 1, 1, 1], dtype=torch.int32, device="cuda")
    info = torch.full((nbatch,), -1, dtype=torch.int32, device="cuda")
    _lib.call("admm_hpd_inverse_batched", n, nbatch, _lib.ptr(d), n * lda, lda, _lib.ptr(work), _lib.ptr(mask), _lib.ptr(info),
              _lib.stream())
    out = d.cpu().numpy()
    inf = info.cpu().numpy()
    for b in range(nbatch):
        if b == 1:
            assert np.array_equal(out[b], buf[b]) and inf[b] == -1
            continue
        if b == 3:
            assert inf[b] != 0 and np.array_equal(out[b], buf[b])
            continue
        ref = np.linalg.inv(G[b])
        err = np.linalg.norm(out[b, :, :n] - ref) / np.linalg.norm(ref)
        assert err < 1e-12, (n, b, err)
        assert inf[b] == 0 and np.array_equal(out[b, :, n:], buf[b, :, n:])
        assert np.abs(out[b, :, :n] @ G[b] - np.eye(n)).max() < 1e-10


@pytest.mark.parametrize("m,n", [(8, 8), (50, 7), (64, 64), (300, 129), (90, 200)])
def test_svd_jacobi_vs_lapack(build_lib, m, n):
    """admm_svd_jacobi (one-sided Jacobi on a cooperative grid) vs np.linalg.svd: singular values, orthogonality and
    reconstruction to working precision; wide matrices (n > m) get n - m zero singular values."""
    from admmsolver_b200 import _dev as D
    rs = np.random.RandomState(m * 1000 + n)
    K = rs.randn(m, n) * np.logspace(0, -6, n)[None, :]            # graded columns
    U, s, V = D.svd_jacobi(torch.from_numpy(K).cuda())
    U, s, V = U.cpu().numpy(), s.cpu().numpy(), V.cpu().numpy()
    sref = np.linalg.svd(K, compute_uv=False)
    r = min(m, n)
    assert np.abs(s[:r] - sref).max() <= 1e-13 * sref[0]
    assert np.all(s[r:] <= 1e-13 * sref[0])
    assert np.abs((U[:, :r] * s[:r]) @ V[:, :r].T - K).max() <= 1e-13 * sref[0]
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    assert np.abs(U[:, :r].T @ U[:, :r] - np.eye(r)).max() < 1e-11        # also for the small singular values


def test_ir_basis_on_device(build_lib):
    """the IR basis generated on the device (admm_svd_jacobi + tensor-core projections) against the host
    construction (np.linalg.svd): same size L = 39 (spm.ipynb:214), singular values to 1e-8 relative (LAPACK itself is
    only accurate to eps * s_0 = 1e-9 s_38 there), the same basis functions, sum rule and sampling matrix; and the SpM
    pipeline built from it -- expansion of the model spectrum, solve, reconstruction (spm.ipynb:155-163,243-300)."""
    from admmsolver_b200 import irbasis, problems
    from admmsolver_b200.batch import SharedSpM
    from oracle import flat
    hb = problems.ir_basis()
    db = irbasis.ir_basis_device()
    assert db.size == hb.size == 39
    s = db.s.cpu().numpy()
    assert np.abs(s / hb.s - 1).max() < 1e-8
    v, u = db.v_omega.cpu().numpy(), db.u_tau.cpu().numpy()
    # orthonormality in the weighted inner products, to working precision for ALL l (Jacobi: relative accuracy)
    assert np.abs((v * hb.womega) @ v.T - np.eye(39)).max() < 1e-12
    assert np.abs((u * hb.wtau) @ u.T - np.eye(39)).max() < 1e-12
    # K = sum_l s_l u_l v_l up to the truncation
    K = problems._kernel(hb.tau, hb.omega, hb.beta)
    Kw = np.sqrt(hb.wtau)[:, None] * K * np.sqrt(hb.womega)[None, :]
    assert np.abs(((u * np.sqrt(hb.wtau)).T * s) @ (v * np.sqrt(hb.womega)) - Kw).max() < 2e-7 * s[0]
    # the leading functions coincide with LAPACK's; the last ones only to the conditioning of their singular value
    # (up to the sign: "largest sample positive" is ambiguous for the odd functions, whose extrema come in +- pairs)
    sgn = np.sign(np.sum((v * hb.womega) * hb.v_omega, axis=1))
    assert np.all(sgn != 0)
    for l in range(39):
        tol = 1e-5 if l < 30 else 1e-2
        assert np.abs(sgn[l] * v[l] - hb.v_omega[l]).max() < tol * np.abs(hb.v_omega[l]).max(), l
    j = np.abs(v).argmax(axis=1)
    assert np.all(v[np.arange(39), j] > 0)                        # the sign convention itself holds
    omega = np.linspace(-10, 10, 400)
    P = db.sampling_matrix(omega).cpu().numpy()
    Ph = (((u * hb.wtau) @ problems._kernel(hb.tau, omega, hb.beta)) / s[:, None]).T
    assert np.abs(P - Ph).max() < 1e-6 * np.abs(Ph).max()
    C = db.sum_rule().cpu().numpy()
    assert np.allclose(C, (v @ hb.womega)[None, :], rtol=1e-10, atol=1e-10)
    # pipeline: rho -> rho_l -> g_l, solve, rho_rec
    rho_l = db.expand_spectrum(problems.rho_three_gaussians(hb.omega)).cpu().numpy()
    assert abs((C @ rho_l)[0] - 1.0) < 1e-5                       # the model spectrum is normalised
    g = -s * rho_l + 1e-4 * np.random.RandomState(0).randn(39)
    eng = SharedSpM(s, P, C, np.array([1.0]), g, lam=1e-4, mu=0.1, batch_wide=True)
    eng.solve(400)
    st = flat.spm_solve(s, P, C, np.array([1.0]), g, 1e-4, 400, mu=0.1)
    x0 = eng.x0()[:, 0]
    assert np.linalg.norm(x0 - st.x0) < 1e-10 * np.linalg.norm(st.x0)
    rec = db.reconstruct(x0.real.copy(), omega).cpu().numpy()
    assert np.abs(rec - P @ x0.real).max() < 1e-12
    # G(tau) -> g_l: the exact G of the model spectrum gives -s_l rho_l
    G = -(K * hb.womega) @ problems.rho_three_gaussians(hb.omega)
    gl = db.project_gtau(G).cpu().numpy()
    assert np.abs(gl - (-s * rho_l)).max() < 1e-9
