"""Objective-function terms with the API of the reference's ``admmsolver.objectivefunc``
(/root/reference/src/admmsolver/objectivefunc.py), evaluated on the GPU.

Each term's contract is the reference's (objectivefunc.py:44-53):
``solve(h, mu)`` returns ``argmin_x F(x) + h^+ x + x^+ h + x^+ mu x``.
``h`` may be a NumPy array (NumPy result, host round trip) or a CUDA tensor (device result, the
path ``SimpleOptimizer`` uses).  Differences from the reference, all behaviour preserving:

* ``alpha A^H y`` is cached (the reference recomputes it on every call, objectivefunc.py:108);
* ``ConstrainedLeastSquares`` caches ``B C^H`` and ``(C B C^H)^-1`` per mu (recomputed every
  iteration in the reference, objectivefunc.py:152-153);
* the latent ``hash(0.0) == 0`` cache collision is not reproduced: the cache
  starts empty.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from . import _dev as D
from .matrix import (DenseMatrix, DiagonalMatrix, MatrixBase, PartialDiagonalMatrix, ScaledIdentityMatrix,
                     asmatrixtype, matrix_hash)

__all__ = ["ObjectiveFunctionBase", "LeastSquares", "ConstrainedLeastSquares", "L1Regularizer", "L2Regularizer",
           "NonNegativePenalty", "SemiPositiveDefinitePenalty"]

Vec = Union[np.ndarray, torch.Tensor]


def _inv_hpd(G: MatrixBase) -> MatrixBase:
    """(alpha A^H A + mu)^-1.  A real dense G of this form is symmetric; when it is also positive definite
    (mu >= 0, the case of every ADMM penalty) the batched tensor-core SPD inverse does it; a complex Hermitian positive
    definite G (complex A) goes through the same kernels in its real form of order 2n; otherwise -- structured or
    indefinite G -- the matrix type's own inv()."""
    if isinstance(G, DenseMatrix):
        g = G._dense_dev()
        if not g.is_complex() and g.shape[0] == g.shape[1] and g.shape[0] <= 512:
            sym = bool(torch.equal(g, g.t())) or float((g - g.t()).abs().max()) <= 1e-13 * float(g.abs().max())
            if sym:
                out = D.spd_inverse(g)
                if out is not None:
                    return DenseMatrix(out)
        elif g.is_complex() and g.shape[0] == g.shape[1] and g.shape[0] <= 256:
            # complex A: alpha A^H A + mu is Hermitian; its real form of order 2n goes through the same kernels
            gh = g.conj().t()
            herm = bool(torch.equal(g, gh)) or float((g - gh).abs().max()) <= 1e-13 * float(g.abs().max())
            if herm:
                out = D.hpd_inverse(g)
                if out is not None:
                    return DenseMatrix(out)
    return G.inv()


def _assert_optional_types(obj, types):
    assert obj is None or isinstance(obj, tuple(types))


def _assert_types(obj, types):
    assert isinstance(obj, tuple(types))


def _ret(x: torch.Tensor, like):
    return x if isinstance(like, torch.Tensor) else D.to_host(x)


class ObjectiveFunctionBase(object):
    """Base class for objective function F(x) (objectivefunc.py:28-53)."""

    def __init__(self, size_x: int) -> None:
        super().__init__()
        self._size_x = size_x

    @property
    def size_x(self) -> int:
        return self._size_x

    def __call__(self, x: np.ndarray) -> float:
        return NotImplemented

    def solve(self, h: Optional[Vec], mu: Optional[MatrixBase]) -> Vec:
        return NotImplemented


class LeastSquares(ObjectiveFunctionBase):
    """alpha * ||y - A @ x||_2^2 (objectivefunc.py:56-110)."""

    def __init__(self, alpha: float, A: Union[np.ndarray, MatrixBase], y: np.ndarray) -> None:
        assert A.ndim == 2
        assert y.ndim == 1
        assert A.shape[0] == y.size
        _assert_types(A, [np.ndarray, torch.Tensor, MatrixBase])
        A = asmatrixtype(A)
        super().__init__(A.shape[1])
        self._alpha = alpha
        self._A = A
        self._y = y
        self._y_dev = D.as_dev(y)
        self._Ac = A.conjugate().T
        self._AcA = self._Ac @ A
        self._Nx = A.shape[1]
        self._Aty = D.axpby(float(alpha), self._Ac @ self._y_dev)        # alpha A^H y, cached
        self._B_cache = (None, None)

    def __call__(self, x: Vec) -> float:
        xd = D.as_dev(x)
        return float(self._alpha * D.sumsq(self._y_dev, self._A @ xd))

    def _get_B(self, mu: MatrixBase) -> MatrixBase:
        """B = (alpha A^H A + mu)^-1, re-inverted only when mu changes (objectivefunc.py:89-96)."""
        hash_ = (type(mu).__name__, matrix_hash(mu))
        if self._B_cache[0] != hash_:
            self._B_cache = (hash_, _inv_hpd((self._alpha * self._AcA) + mu))
            self._on_new_B()
        return self._B_cache[1]

    def _on_new_B(self) -> None:
        pass

    def _default_mu(self) -> MatrixBase:
        return ScaledIdentityMatrix(self._Nx, 0.0)

    def _rhs(self, h: Optional[Vec]) -> torch.Tensor:
        if h is None:
            return self._Aty
        hd = D.as_dev(h)
        assert tuple(hd.shape) == (self._Nx,)
        return D.axpby(1.0, self._Aty, -1.0, hd)

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None) -> Vec:
        _assert_optional_types(h, [np.ndarray, torch.Tensor])
        _assert_optional_types(mu, [MatrixBase])
        if mu is None:
            mu = self._default_mu()
        assert tuple(mu.shape) == (self._Nx, self._Nx)
        return _ret(self._get_B(mu) @ self._rhs(h), h)


class ConstrainedLeastSquares(LeastSquares):
    r"""alpha * ||y - A @ x||_2^2 subject to C @ x = D (objectivefunc.py:113-157)."""

    def __init__(self, alpha: float, A, y: np.ndarray, C, D_: np.ndarray) -> None:
        assert A.ndim == 2
        assert y.ndim == 1
        assert C.ndim == 2
        assert D_.ndim == 1
        assert A.shape[0] == y.size
        assert A.shape[1] == C.shape[1]
        assert C.shape[0] == D_.size
        _assert_types(C, [np.ndarray, torch.Tensor, MatrixBase])
        super().__init__(alpha, asmatrixtype(A), y)
        self._C = asmatrixtype(C)
        self._D = D_
        self._D_dev = D.as_dev(D_)
        self._Ch = self._C.conjugate().T
        self._xi2 = None
        self._CBCinv = None

    def _on_new_B(self) -> None:
        B = self._B_cache[1]
        self._xi2 = -(B @ self._Ch)                       # xi2 = -B C^H
        self._CBCinv = (self._C @ self._xi2).inv()        # (C xi2)^-1

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None) -> Vec:
        _assert_optional_types(h, [np.ndarray, torch.Tensor])
        _assert_optional_types(mu, [MatrixBase])
        if mu is None:
            mu = self._default_mu()
        assert tuple(mu.shape) == (self._Nx, self._Nx)
        B = self._get_B(mu)
        xi1 = B @ self._rhs(h)
        tmp2 = D.axpby(1.0, self._D_dev, -1.0, self._C @ xi1)
        nu = self._CBCinv @ tmp2
        return _ret(D.axpby(1.0, xi1, 1.0, self._xi2 @ nu), h)


def _mu_diag_dev(mu: MatrixBase) -> torch.Tensor:
    assert isinstance(mu, (DiagonalMatrix, ScaledIdentityMatrix))
    d = mu._diag_dev()
    if d.is_complex():
        d = d.real.contiguous()
    return d


class L1Regularizer(ObjectiveFunctionBase):
    """F(x) = alpha * |x|_1 (objectivefunc.py:160-195)."""

    def __init__(self, alpha: float, size_x: int) -> None:
        assert isinstance(size_x, int), type(size_x)
        super().__init__(size_x)
        assert alpha > 0
        self._alpha = alpha

    def __call__(self, x) -> float:
        return float(self._alpha * np.sum(np.abs(D.to_host(x) if isinstance(x, torch.Tensor) else x)))

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None) -> Vec:
        """soft(-Re(h)/mu, 0.5 alpha/mu); needs a diagonal mu; returns a real vector."""
        _assert_types(h, [np.ndarray, torch.Tensor])
        assert isinstance(mu, DiagonalMatrix) or isinstance(mu, ScaledIdentityMatrix)
        hd = D.as_dev(h)
        return _ret(D.prox_l1(hd, _mu_diag_dev(mu), self._alpha, complex_out=False), h)

    def _solve_complex(self, h: torch.Tensor, mu: MatrixBase) -> torch.Tensor:
        return D.prox_l1(h, _mu_diag_dev(mu), self._alpha, complex_out=True)


class L2Regularizer(ObjectiveFunctionBase):
    """F(x) = alpha * |A x|_2^2 (objectivefunc.py:198-242).  Row f1 of SURVEY.md 8(f): same cached
    inverse as LeastSquares with A^H y = 0."""

    def __init__(self, alpha: float, A: Union[np.ndarray, MatrixBase]):
        _assert_optional_types(A, [np.ndarray, torch.Tensor, MatrixBase])
        A = asmatrixtype(A)
        super().__init__(A.shape[1])
        assert alpha > 0
        self._alpha = alpha
        self._A = A
        self._AcA = A.conjugate().T @ A
        self._B_cache = (None, None)

    def __call__(self, x: Vec):
        return float(self._alpha * D.sumsq(self._A @ D.as_dev(x)))

    def _get_B(self, mu: MatrixBase):
        hash_ = (type(mu).__name__, matrix_hash(mu))
        if self._B_cache[0] != hash_:
            self._B_cache = (hash_, _inv_hpd((self._alpha * self._AcA) + mu))
        return self._B_cache[1]

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None):
        _assert_optional_types(h, [np.ndarray, torch.Tensor])
        _assert_optional_types(mu, [MatrixBase])
        if mu is None:
            mu = ScaledIdentityMatrix(self._A.shape[1], 0.0)
        if h is None:
            return np.zeros(self._A.shape[1])
        hd = D.as_dev(h)
        return _ret(D.axpby(-1.0, self._get_B(mu) @ hd), h)


class NonNegativePenalty(ObjectiveFunctionBase):
    """F(x) = infty * Theta(-x) (objectivefunc.py:245-271)."""

    def __init__(self, size_x: int):
        super().__init__(size_x)

    def __call__(self, x):
        return 0.0

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None):
        assert isinstance(h, (np.ndarray, torch.Tensor))
        assert isinstance(mu, DiagonalMatrix) or isinstance(mu, ScaledIdentityMatrix)
        hd = D.as_dev(h)
        return _ret(D.prox_nonneg(hd, _mu_diag_dev(mu), complex_out=False), h)

    def _solve_complex(self, h: torch.Tensor, mu: MatrixBase) -> torch.Tensor:
        return D.prox_nonneg(h, _mu_diag_dev(mu), complex_out=True)


class SemiPositiveDefinitePenalty(ObjectiveFunctionBase):
    """Penalty for negative eigenvalues (objectivefunc.py:274-327): x is reshaped to a three-way tensor
    and every slice along ``axis`` is projected onto the PSD cone.  SURVEY.md 8(f) row f2: the per-slice
    ``np.linalg.eigh`` loop of the reference is one batched Jacobi kernel (``admm_prox_psd``, one warp
    per slice, slices up to 32 x 32)."""

    def __init__(self, shape, axis: int):
        assert len(shape) == 3
        super().__init__(int(np.prod(shape)))
        self._shape = tuple(int(v) for v in shape)
        self._axis = axis

    def __call__(self, x):
        return 0.0

    def _diagonals(self, mu: MatrixBase) -> torch.Tensor:
        """mu as a vector of diagonal entries (objectivefunc.py:296-311)."""
        assert isinstance(mu, (DiagonalMatrix, ScaledIdentityMatrix)) or \
            (isinstance(mu, PartialDiagonalMatrix) and isinstance(mu.matrix, (ScaledIdentityMatrix, DiagonalMatrix)))
        if isinstance(mu, (DiagonalMatrix, ScaledIdentityMatrix)):
            d = mu._diag_dev()
        else:
            inner = mu.matrix._diag_dev()
            rest = int(np.prod(mu.rest_dims))
            d = inner.reshape(-1, 1).expand(inner.numel(), rest).reshape(-1)      # einsum('i,j->ij', diag, ones).ravel()
        if d.is_complex():
            d = d.real
        d = d.contiguous()
        assert d.ndim == 1 and d.numel() == self.size_x
        return d

    def solve(self, h: Optional[Vec] = None, mu: Optional[MatrixBase] = None):
        """Works only if mu is (partially) diagonal; the imaginary part of h is dropped like in the reference."""
        assert isinstance(h, (np.ndarray, torch.Tensor))
        if mu is None:
            raise ValueError("mu must not be None!")
        hd = D.as_dev(h)
        return _ret(D.prox_psd(hd, self._diagonals(mu), self._shape, self._axis, complex_out=False), h)

    def _solve_complex(self, h: torch.Tensor, mu: MatrixBase) -> torch.Tensor:
        return D.prox_psd(h, self._diagonals(mu), self._shape, self._axis, complex_out=True)
