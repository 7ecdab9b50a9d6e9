"""Thin functional layer over the C ABI primitives: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is plumbing here (allocation, dtype views, host<->device copies); every arithmetic
operation is a kernel of ``libadmm_b200.so``.
"""
from __future__ import annotations

import numpy as np
import torch

import ctypes as C

from . import _lib
from ._lib import call, ptr, stream

F64 = torch.float64
C128 = torch.complex128


def device() -> torch.device:
    return _lib.require_cuda()


def as_dev(a, dtype=None) -> torch.Tensor:
    """NumPy array / scalar sequence / tensor -> contiguous float64 or complex128 CUDA tensor."""
    dev = device()
    if isinstance(a, torch.Tensor):
        t = a.to(dev)
    else:
        arr = np.asarray(a)
        if arr.dtype.kind == "c":
            arr = arr.astype(np.complex128, copy=False)
        else:
            arr = arr.astype(np.float64, copy=False)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    if t.dtype not in (F64, C128):
        t = t.to(C128 if t.is_complex() else F64)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def ncomp(t: torch.Tensor) -> int:
    return 2 if t.is_complex() else 1


def gemm(op: int, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """op(A) @ B with B of shape (k,) or (k, n); mixed real/complex operands are promoted."""
    vec = B.ndim == 1
    B2 = B.reshape(B.shape[0], -1).contiguous()
    A = A.contiguous()
    m = A.shape[0] if op == _lib.OP_N else A.shape[1]
    k = A.shape[1] if op == _lib.OP_N else A.shape[0]
    assert B2.shape[0] == k, f"shape mismatch {tuple(A.shape)} op={op} @ {tuple(B.shape)}"
    n = B2.shape[1]
    if not A.is_complex() and B2.is_complex():
        # real operator on complex data: one real GEMM over the interleaved (re, im) columns
        Br = torch.view_as_real(B2).reshape(k, 2 * n)
        Cr = torch.empty(m, 2 * n, dtype=F64, device=A.device)
        call("admm_gemm", 0, op, m, 2 * n, k, ptr(A), A.shape[1], ptr(Br), 2 * n, ptr(Cr), 2 * n, stream())
        out = torch.view_as_complex(Cr.reshape(m, n, 2))
    else:
        if A.is_complex() and not B2.is_complex():
            B2 = B2.to(C128)
        cplx = A.is_complex()
        out = torch.empty(m, n, dtype=C128 if cplx else F64, device=A.device)
        call("admm_gemm", int(cplx), op, m, n, k, ptr(A), A.shape[1], ptr(B2), n, ptr(out), n, stream())
    return out.reshape(m) if vec else out.reshape((m,) + tuple(B.shape[1:]))


def diag_mul(d: torch.Tensor, V: torch.Tensor, rows_out: int) -> torch.Tensor:
    """(rectangular diagonal with entries d, ``rows_out`` rows) @ V along the first axis of V."""
    vec = V.ndim == 1
    V2 = V.reshape(V.shape[0], -1).contiguous()
    cplx = d.is_complex() or V2.is_complex()
    if cplx:
        d, V2 = d.to(C128), V2.to(C128)
    nd = min(d.numel(), V2.shape[0])
    n = V2.shape[1]
    out = torch.empty(rows_out, n, dtype=C128 if cplx else F64, device=V2.device)
    call("admm_diag_mul", int(cplx), rows_out, min(nd, rows_out), n, ptr(d.contiguous()), ptr(V2), n, ptr(out), n, stream())
    return out.reshape(rows_out) if vec else out.reshape((rows_out,) + tuple(V.shape[1:]))


def axpby(a: float, x: torch.Tensor, b: float = 0.0, y: torch.Tensor = None) -> torch.Tensor:
    """a*x + b*y with real scalars (complex tensors are handled as interleaved doubles)."""
    if y is not None and x.is_complex() != y.is_complex():
        x, y = x.to(C128), y.to(C128)
    x = x.contiguous()
    out = torch.empty_like(x)
    n = x.numel() * ncomp(x)
    call("admm_axpby", n, float(a), ptr(x), float(b), ptr(y.contiguous()) if y is not None else None, ptr(out), stream())
    return out


def scale(c, x: torch.Tensor) -> torch.Tensor:
    """c*x for a real or complex scalar c."""
    c = complex(c) if isinstance(c, (complex, np.complexfloating)) else float(c)
    if isinstance(c, complex):
        if c.imag == 0.0:
            return axpby(c.real, x.to(C128))
        d = torch.full((1,), c, dtype=C128, device=x.device).expand(x.shape[0] if x.ndim else 1).contiguous()
        return diag_mul(d, x.to(C128), x.shape[0])
    return axpby(c, x)


def unary(op: int, x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    out = torch.empty_like(x)
    call("admm_ewise_unary", op, int(x.is_complex()), x.numel(), ptr(x), ptr(out), stream())
    return out


def recip(x):
    return unary(0, x)


def conj(x):
    return unary(1, x) if x.is_complex() else x


_scratch = {}


def sumsq(x: torch.Tensor, y: torch.Tensor = None) -> float:
    """||x||^2 or ||x - y||^2 (synchronises: the value is returned to the host)."""
    if y is not None and x.is_complex() != y.is_complex():
        x, y = x.to(C128), y.to(C128)
    x = x.contiguous()
    dev = x.device
    if dev not in _scratch:
        _scratch[dev] = (torch.zeros(1024, dtype=F64, device=dev), torch.zeros(1, dtype=F64, device=dev))
    scratch, out = _scratch[dev]
    call("admm_sumsq", x.numel() * ncomp(x), ptr(x), ptr(y.contiguous()) if y is not None else None, ptr(out),
         ptr(scratch), stream())
    return float(out.item())


def sumsq_into(out: torch.Tensor, slot: int, x: torch.Tensor, y: torch.Tensor = None) -> None:
    """out[slot] = ||x||^2 or ||x - y||^2 without a host synchronisation (stream ordered)."""
    if y is not None and x.is_complex() != y.is_complex():
        x, y = x.to(C128), y.to(C128)
    x = x.contiguous()
    dev = x.device
    if dev not in _scratch:
        _scratch[dev] = (torch.zeros(1024, dtype=F64, device=dev), torch.zeros(1, dtype=F64, device=dev))
    scratch = _scratch[dev][0]
    call("admm_sumsq", x.numel() * ncomp(x), ptr(x), ptr(y.contiguous()) if y is not None else None,
         ptr(out[slot:slot + 1]), ptr(scratch), stream())


def pair_norms_into(out: torch.Tensor, slot: int, p1, p2, d1, d2) -> None:
    """out[slot:slot+6] = |p1-p2|^2, |p1|^2, |p2|^2, |d1-d2|^2, |d1|^2, |d2|^2 in one pass (two launches instead
    of twelve), stream ordered."""
    def same(a, b):
        if a.is_complex() != b.is_complex():
            a, b = a.to(C128), b.to(C128)
        return a.contiguous(), b.contiguous()
    p1, p2 = same(p1, p2)
    d1, d2 = same(d1, d2)
    assert p1.numel() == p2.numel() and d1.numel() == d2.numel()
    dev = p1.device
    if dev not in _scratch6:
        _scratch6[dev] = torch.zeros(6 * 1024, dtype=F64, device=dev)
    call("admm_pair_norms", p1.numel() * ncomp(p1), ptr(p1), ptr(p2), d1.numel() * ncomp(d1), ptr(d1), ptr(d2),
         ptr(out[slot:slot + 6]), ptr(_scratch6[dev]), stream())


_scratch6 = {}


def norm(x: torch.Tensor, y: torch.Tensor = None) -> float:
    return float(np.sqrt(sumsq(x, y)))


def spd_inverse(A: torch.Tensor):
    """Inverse of a real symmetric positive definite matrix on the tensor cores (n <= 512 stays on the
    DMMA kernels); returns None when a non-positive pivot shows up (caller falls back to `inverse`)."""
    n = A.shape[0]
    out = A.contiguous().clone()
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    call("admm_spd_inverse_batched", n, 1, ptr(out), n * n, n, None, ptr(info), stream())
    if int(info.item()) != 0:
        return None
    return out


def hpd_inverse(A: torch.Tensor):
    """Inverse of a complex Hermitian positive definite matrix on the tensor cores (real form of order 2n through the
    batched SPD kernels); returns None when a non-positive pivot shows up (caller falls back to `inverse`)."""
    n = A.shape[0]
    out = A.contiguous().clone()
    work = torch.empty(4 * n * n, dtype=F64, device=A.device)
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    call("admm_hpd_inverse_batched", n, 1, ptr(out), n * n, n, ptr(work), None, ptr(info), stream())
    if int(info.item()) != 0:
        return None
    return out


def inverse(A: torch.Tensor) -> torch.Tensor:
    """General inverse (Gauss-Jordan with partial pivoting on the device)."""
    n = A.shape[0]
    assert A.shape == (n, n)
    A = A.contiguous()
    out = torch.empty_like(A)
    work = torch.empty(n, 2 * n, dtype=A.dtype, device=A.device)
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    call("admm_inverse", int(A.is_complex()), n, ptr(A), n, ptr(out), n, ptr(work), ptr(info), stream())
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("Singular matrix")
    return out


def prox_l1(h: torch.Tensor, mu_diag: torch.Tensor, alpha: float, complex_out: bool) -> torch.Tensor:
    n = h.numel()
    h = h.contiguous()
    out = torch.empty(n, dtype=C128 if complex_out else F64, device=h.device)
    call("admm_prox_l1", n, ptr(h), ncomp(h), ptr(mu_diag.contiguous()), float(alpha), ptr(out),
         2 if complex_out else 1, stream())
    return out


def prox_psd(h: torch.Tensor, mu_diag: torch.Tensor, shape, axis: int, complex_out: bool) -> torch.Tensor:
    """Per-slice PSD projection of -Re(h)/mu reshaped to the 3-way ``shape`` (C order); the matrices are
    the slices along ``axis`` (rows / columns = the two remaining axes in order)."""
    s0, s1, s2 = (int(v) for v in shape)
    strides = (s1 * s2, s2, 1)
    rest = [a for a in range(3) if a != axis % 3]
    dims = (s0, s1, s2)
    n = dims[rest[0]]
    assert dims[rest[1]] == n, "SemiPositiveDefinitePenalty needs square slices"
    nb = dims[axis % 3]
    total = s0 * s1 * s2
    h = h.contiguous()
    assert h.numel() == total and mu_diag.numel() == total
    out = torch.empty(total, dtype=C128 if complex_out else F64, device=h.device)
    grid = C.c_longlong(0)
    work = None
    if _lib.lib.admm_prox_psd_work_doubles(n, nb, C.byref(grid)) != 0:       # slices beyond 160 x 160: L2-resident workspace
        work = torch.empty(int(grid.value) * n * n, dtype=F64, device=h.device)
    call("admm_prox_psd", n, nb, strides[axis % 3], strides[rest[0]], strides[rest[1]], ptr(h), ncomp(h),
         ptr(mu_diag.contiguous()), ptr(out), 2 if complex_out else 1, ptr(work), stream())
    return out


def svd_jacobi(K: torch.Tensor, max_sweeps: int = 40, return_sweeps: bool = False):
    """SVD of a real (m, n) device matrix by one-sided Jacobi (``admm_svd_jacobi``): returns ``U`` (m, n), ``s`` (n,)
    descending and ``V`` (n, n) with ``K = U diag(s) V^T``; columns of ``U`` belonging to singular values at the
    rounding level of the largest one are unit vectors of noise, as with LAPACK."""
    m, n = K.shape
    Wt = K.t().contiguous().to(F64)                      # rows = columns of K
    Vt = torch.eye(n, dtype=F64, device=K.device)
    sv = torch.empty(n, dtype=F64, device=K.device)
    scratch = torch.zeros(max_sweeps, dtype=F64, device=K.device)
    info = torch.zeros(1, dtype=torch.int32, device=K.device)
    call("admm_svd_jacobi", m, n, ptr(Wt), m, ptr(Vt), n, ptr(sv), ptr(scratch), int(max_sweeps), ptr(info), stream())
    if int(info.item()) < 0:
        raise np.linalg.LinAlgError("admm_svd_jacobi did not converge in %d sweeps" % max_sweeps)
    order = torch.argsort(sv, descending=True)
    s_sorted = sv[order]
    U = (Wt[order] / s_sorted.clamp_min(1e-300)[:, None]).t().contiguous()
    V = Vt[order].t().contiguous()
    if return_sweeps:
        return U, s_sorted, V, int(info.item())
    return U, s_sorted, V


def prox_nonneg(h: torch.Tensor, mu_diag: torch.Tensor, complex_out: bool) -> torch.Tensor:
    n = h.numel()
    h = h.contiguous()
    out = torch.empty(n, dtype=C128 if complex_out else F64, device=h.device)
    call("admm_prox_nonneg", n, ptr(h), ncomp(h), ptr(mu_diag.contiguous()), ptr(out), 2 if complex_out else 1, stream())
    return out
