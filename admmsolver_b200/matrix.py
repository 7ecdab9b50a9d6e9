"""Structured-matrix types with the API of the reference's ``admmsolver.matrix``
(/root/reference/src/admmsolver/matrix.py), backed by CUDA.

Same names, constructor arguments, attributes (``shape ndim data diagonals coeff matrix
rest_dims``), operators (``@ * + - neg T conjugate()/conj() inv() hash() asmatrix() is_diagonal()``)
and error behaviour as the reference; the difference is where the numbers live and who does the
arithmetic: operands are mirrored to the device on first use and every product, sum, inverse and
mat-vec is a kernel of ``libadmm_b200.so`` (see ``_dev.py``).  ``matrix @ ndarray`` takes and
returns NumPy arrays (host round trip, drop-in behaviour); ``matrix @ torch.cuda tensor`` stays on
the device -- that is the path the optimizer uses.

Structure-preserving dispatch follows the reference line by line in *behaviour* (which result type
each pair of operand types yields -- matrix.py:100-118, 255-295, 342-354, 453-513), not in code.

The class / method names and the name-based ``_add_X_Y`` dispatch skeleton are the reference's public
API and are therefore derived from it:
    SPDX-License-Identifier: MIT
    admmsolver -- Copyright (c) 2021- Hiroshi Shinaoka and others (LICENSE.txt of SpM-lab/admmsolver)
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _dev as D
from ._lib import OP_H, OP_N, OP_T

__all__ = ["MatrixBase", "DenseMatrix", "ScaledIdentityMatrix", "DiagonalMatrix", "PartialDiagonalMatrix",
           "identity", "matrix_hash", "asmatrixtype"]

_SCALARS = (complex, float, np.float64, np.complex128)
ArrayLike = Union[np.ndarray, torch.Tensor]


def _is_array(x) -> bool:
    return isinstance(x, (np.ndarray, torch.Tensor))


def _wrap_like(result: torch.Tensor, like):
    """Return ``result`` as NumPy when the caller passed NumPy, else leave it on the device."""
    return D.to_host(result) if isinstance(like, np.ndarray) else result


class MatrixBase(object):
    """Protocol of matrix.py:9-60."""

    def __init__(self) -> None:
        super().__init__()
        self.shape = (1, 1)
        self.ndim = 2

    def is_diagonal(self) -> bool:
        return self.shape[0] == self.shape[1]

    def __neg__(self) -> "MatrixBase":
        return -1.0 * self

    def asmatrix(self) -> np.ndarray:
        raise NotImplementedError

    def _dense_dev(self) -> torch.Tensor:
        """Dense (m x n) device tensor of this operator."""
        raise NotImplementedError

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        """self @ v for a device vector (n,) or stack (n, k): the optimizer's hot call."""
        raise NotImplementedError

    def __mul__(self, other):
        return NotImplemented

    __rmul__ = __mul__

    def __matmul__(self, other):
        return NotImplemented

    def __add__(self, other: "MatrixBase") -> "MatrixBase":
        return _add_matrices(self, other)

    def conjugate(self) -> "MatrixBase":
        raise NotImplementedError

    def conj(self) -> "MatrixBase":
        return self.conjugate()

    @property
    def T(self) -> "MatrixBase":
        raise NotImplementedError

    def __sub__(self, other) -> "MatrixBase":
        return self + (-other)

    def inv(self) -> "MatrixBase":
        raise NotImplementedError

    def hash(self) -> int:
        raise NotImplementedError

    def _matmul_array(self, other: ArrayLike):
        assert self.shape[1] == other.shape[0], f"{self.shape} {tuple(other.shape)}"
        v = D.as_dev(other)
        return _wrap_like(self._apply(v), other)


# ------------------------------------------------------------------------------------------------
class DenseMatrix(MatrixBase):
    """ndarray wrapper (matrix.py:63-121)."""

    def __init__(self, matrix: ArrayLike) -> None:
        assert _is_array(matrix)
        assert matrix.ndim == 2
        self._host = matrix if isinstance(matrix, np.ndarray) else None
        self._t = matrix if isinstance(matrix, torch.Tensor) else None
        self.shape = (int(matrix.shape[0]), int(matrix.shape[1]))
        self.ndim = 2
        self._hash = None

    @property
    def data(self) -> np.ndarray:
        if self._host is None:
            self._host = D.to_host(self._t)
        return self._host

    def _dense_dev(self) -> torch.Tensor:
        if self._t is None:
            self._t = D.as_dev(self._host)
        elif self._t.dtype not in (D.F64, D.C128) or not self._t.is_cuda:
            self._t = D.as_dev(self._t)
        return self._t

    def hash(self) -> int:
        if self._hash is None:
            self._hash = matrix_hash(self.asmatrix())
        return self._hash

    def asmatrix(self) -> np.ndarray:
        return self.data

    def inv(self) -> "DenseMatrix":
        return DenseMatrix(D.inverse(self._dense_dev()))

    @property
    def T(self) -> "DenseMatrix":
        return DenseMatrix(self._dense_dev().t().contiguous())

    def conjugate(self) -> "DenseMatrix":
        return DenseMatrix(D.conj(self._dense_dev()))

    conj = conjugate

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        return D.gemm(OP_N, self._dense_dev(), v)

    def _apply_adjoint(self, v: torch.Tensor) -> torch.Tensor:
        """self^H @ v without materialising the adjoint (the reference rebuilds it every call,
        optimizer.py:187,198)."""
        return D.gemm(OP_H, self._dense_dev(), v)

    def __mul__(self, other) -> "DenseMatrix":
        if np.isscalar(other):
            return DenseMatrix(D.scale(other, self._dense_dev()))
        return NotImplemented

    __rmul__ = __mul__

    def __matmul__(self, other):
        assert self.shape[1] == other.shape[0]
        assert isinstance(other, MatrixBase) or (_is_array(other) and other.ndim <= 2)
        if _is_array(other):
            return self._matmul_array(other)
        if isinstance(other, ScaledIdentityMatrix):
            return self @ other.to_diagonal_matrix()
        if isinstance(other, DiagonalMatrix):
            # column scaling, zero padded to (m, other.shape[1]); complex128 like matrix.py:111-116
            m, ncol = self.shape[0], other.shape[1]
            mn = min(*other.shape)
            At = self._dense_dev()[:, :mn].t().contiguous()             # (mn, m)
            scaled = D.diag_mul(other._diag_dev(), At, ncol)             # (ncol, m), zero rows beyond mn
            return DenseMatrix(scaled.t().contiguous().to(D.C128))
        return DenseMatrix(D.gemm(OP_N, self._dense_dev(), other._dense_dev()))


# ------------------------------------------------------------------------------------------------
class ScaledIdentityMatrix(MatrixBase):
    """``coeff * I`` with a possibly rectangular shape (matrix.py:124-194)."""

    def __init__(self, shape: Union[int, Tuple[int, int]], coeff) -> None:
        assert type(coeff) in [complex, float, np.float64, np.complex128], type(coeff)
        self.shape = (0, 0)
        if isinstance(shape, (int, np.integer)) and not isinstance(shape, bool):
            self.shape = (int(shape), int(shape))
        elif isinstance(shape, tuple):
            self.shape = (int(shape[0]), int(shape[1]))
        else:
            raise ValueError("Invalid shape value!")
        self.coeff = coeff
        self.ndim = 2

    def hash(self) -> int:
        return matrix_hash(self.coeff)

    def asmatrix(self) -> np.ndarray:
        return self.coeff * np.eye(N=self.shape[0], M=self.shape[1])

    def _dense_dev(self) -> torch.Tensor:
        return self.to_diagonal_matrix()._dense_dev()

    def inv(self) -> "ScaledIdentityMatrix":
        if not self.is_diagonal():
            raise RuntimeError("A rectangular matrix is not invertible!")
        return ScaledIdentityMatrix(self.shape, 1 / self.coeff)

    @property
    def T(self) -> "ScaledIdentityMatrix":
        return ScaledIdentityMatrix((self.shape[1], self.shape[0]), self.coeff)

    @property
    def diagonals(self) -> np.ndarray:
        if not self.is_diagonal():
            raise RuntimeError("Diagonals of a rectangular matrix is ill defined!")
        return np.full(self.shape[0], self.coeff)

    def _diag_dev(self) -> torch.Tensor:
        n = min(*self.shape)
        cplx = isinstance(self.coeff, (complex, np.complexfloating))
        return torch.full((n,), complex(self.coeff) if cplx else float(self.coeff),
                          dtype=D.C128 if cplx else D.F64, device=D.device())

    def conjugate(self) -> "ScaledIdentityMatrix":
        return ScaledIdentityMatrix(self.shape, np.conjugate(self.coeff))

    conj = conjugate

    def __mul__(self, other) -> "ScaledIdentityMatrix":
        if type(other) in [complex, float, np.float64, np.complex128]:
            return ScaledIdentityMatrix(self.shape, self.coeff * other)
        return NotImplemented

    __rmul__ = __mul__

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        if self.is_diagonal():
            return D.scale(self.coeff, v)
        return D.diag_mul(self._diag_dev(), v, self.shape[0])

    def _apply_adjoint(self, v: torch.Tensor) -> torch.Tensor:
        return self.T.conjugate()._apply(v)

    def __matmul__(self, other):
        assert self.shape[1] == other.shape[0], f"{self.shape} {other.shape}"
        assert isinstance(other, MatrixBase) or _is_array(other)
        if _is_array(other):
            return self._matmul_array(other)
        return self.to_diagonal_matrix() @ other

    def to_diagonal_matrix(self) -> "DiagonalMatrix":
        return DiagonalMatrix(self._diag_dev(), self.shape)


# ------------------------------------------------------------------------------------------------
class DiagonalMatrix(MatrixBase):
    """Diagonal matrix, also rectangular: zero padded / truncated (matrix.py:197-298)."""

    def __init__(self, diagonals: ArrayLike, shape: Optional[Tuple[int, int]] = None) -> None:
        assert diagonals.ndim == 1
        self._host = diagonals if isinstance(diagonals, np.ndarray) else None
        self._t = diagonals if isinstance(diagonals, torch.Tensor) else None
        self.ndim = 2
        size = int(diagonals.shape[0])
        if shape is None:
            self.shape = (size, size)
        else:
            self.shape = (int(shape[0]), int(shape[1]))
        assert min(*self.shape) == size, f"{self.shape} {size}"
        self._hash = None

    def hash(self) -> int:
        if self._hash is None:
            self._hash = matrix_hash(self.diagonals)
        return self._hash

    @property
    def diagonals(self) -> np.ndarray:
        if self._host is None:
            self._host = D.to_host(self._t)
        return self._host

    @property
    def _diagonals(self) -> np.ndarray:
        return self.diagonals

    def _diag_dev(self) -> torch.Tensor:
        if self._t is None:
            self._t = D.as_dev(self._host)
        elif self._t.dtype not in (D.F64, D.C128) or not self._t.is_cuda:
            self._t = D.as_dev(self._t)
        return self._t

    def inv(self) -> "DiagonalMatrix":
        if not self.is_diagonal():
            raise RuntimeError("Must be a diagonal matrix!")
        return DiagonalMatrix(D.recip(self._diag_dev()))

    def asmatrix(self) -> np.ndarray:
        return D.to_host(self._dense_dev())

    def _dense_dev(self) -> torch.Tensor:
        d = self._diag_dev()
        n = d.numel()
        eye = torch.zeros(n, self.shape[1], dtype=d.dtype, device=d.device)
        eye[:, :n].fill_diagonal_(1.0)                   # structural ones (no arithmetic)
        return D.diag_mul(d, eye, self.shape[0])

    @property
    def T(self) -> "DiagonalMatrix":
        return DiagonalMatrix(self._diag_dev(), shape=(self.shape[1], self.shape[0]))

    def conjugate(self) -> "DiagonalMatrix":
        return DiagonalMatrix(D.conj(self._diag_dev()), self.shape)

    conj = conjugate

    def __mul__(self, other) -> "DiagonalMatrix":
        if type(other) in [complex, float, np.float64, np.complex128]:
            return DiagonalMatrix(D.scale(other, self._diag_dev()), self.shape)
        return NotImplemented

    __rmul__ = __mul__

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        return D.diag_mul(self._diag_dev(), v, self.shape[0])

    def _apply_adjoint(self, v: torch.Tensor) -> torch.Tensor:
        return D.diag_mul(D.conj(self._diag_dev()), v, self.shape[1])

    def __matmul__(self, other):
        assert self.shape[1] == other.shape[0]
        assert isinstance(other, MatrixBase) or _is_array(other)
        if _is_array(other):
            return self._matmul_array(other)
        if isinstance(other, DenseMatrix):
            return DenseMatrix(D.diag_mul(self._diag_dev(), other._dense_dev(), self.shape[0]))
        if isinstance(other, DiagonalMatrix):
            min_size = min(self.shape[0], other.shape[1])
            prod = D.diag_mul(self._diag_dev(), other._diag_dev(), min_size)     # zero padded product
            return DiagonalMatrix(prod, (self.shape[0], other.shape[1]))
        if isinstance(other, PartialDiagonalMatrix):
            diags = self.diagonals.reshape(other.matrix.shape[0], -1)
            if np.allclose(diags, diags[:, 0:1]):
                inner = DiagonalMatrix(np.ascontiguousarray(diags[:, 0])) @ DenseMatrix(other.matrix._dense_dev())
                return PartialDiagonalMatrix(inner, other.rest_dims)
            return DenseMatrix(D.diag_mul(self._diag_dev(), other._dense_dev(), self.shape[0]))
        if isinstance(other, ScaledIdentityMatrix):
            return self @ other.to_diagonal_matrix()
        return NotImplemented

    def __str__(self) -> str:
        return "DiagonalMatrix: " + self.diagonals.__str__()


# ------------------------------------------------------------------------------------------------
class PartialDiagonalMatrix(MatrixBase):
    """``A (x) I_rest``: the reference's batching device (matrix.py:301-401).  Vector index =
    row * prod(rest) + batch, i.e. the batch index is the fastest one."""

    def __init__(self, matrix: Union[ArrayLike, MatrixBase], rest_dims: tuple) -> None:
        assert matrix.ndim == 2
        self.matrix = asmatrixtype(matrix)
        self.rest_dims = rest_dims
        self.ndim = 2
        nrest = int(np.prod(rest_dims))
        self.shape = (self.matrix.shape[0] * nrest, self.matrix.shape[1] * nrest)

    def hash(self) -> int:
        return matrix_hash(self.matrix)

    def asmatrix(self) -> np.ndarray:
        return D.to_host(self._dense_dev())

    def _dense_dev(self) -> torch.Tensor:
        nrest = int(np.prod(self.rest_dims))
        eye = torch.eye(self.shape[1], dtype=D.F64, device=D.device())     # structural
        return self._apply(eye)

    def inv(self) -> "PartialDiagonalMatrix":
        return PartialDiagonalMatrix(self.matrix.inv(), self.rest_dims)

    @property
    def T(self) -> "PartialDiagonalMatrix":
        return PartialDiagonalMatrix(self.matrix.T, self.rest_dims)

    def conjugate(self) -> "PartialDiagonalMatrix":
        return PartialDiagonalMatrix(self.matrix.conjugate(), self.rest_dims)

    conj = conjugate

    def _apply(self, v: torch.Tensor) -> torch.Tensor:
        return _matvec_impl(self.matrix, v, self.rest_dims)

    def _apply_adjoint(self, v: torch.Tensor) -> torch.Tensor:
        return _matvec_impl(self.matrix, v, self.rest_dims, adjoint=True)

    def __matmul__(self, other):
        assert self.shape[1] == other.shape[0]
        assert isinstance(other, MatrixBase) or _is_array(other)
        if _is_array(other):
            return self.matvec(other)
        if isinstance(other, PartialDiagonalMatrix) and self.rest_dims == other.rest_dims:
            return PartialDiagonalMatrix(self.matrix @ other.matrix, self.rest_dims)
        if isinstance(other, ScaledIdentityMatrix) and other.is_diagonal():
            return PartialDiagonalMatrix(other.coeff * self.matrix, self.rest_dims)
        return DenseMatrix(D.gemm(OP_N, self._dense_dev(), other._dense_dev()))

    def __mul__(self, other) -> "PartialDiagonalMatrix":
        if type(other) in [float, complex, np.float64, np.complex128]:
            return PartialDiagonalMatrix(self.matrix * other, self.rest_dims)
        return NotImplemented

    __rmul__ = __mul__

    def matvec(self, v: ArrayLike):
        r"""(a \otimes I) @ v for a vector or a stack of vectors (first axis)."""
        return _wrap_like(self._apply(D.as_dev(v)), v)


def _matvec_impl(matrix: MatrixBase, v: torch.Tensor, rest_dims: tuple, adjoint: bool = False) -> torch.Tensor:
    """Apply ``matrix (x) I`` (or its adjoint) to a device vector / stack (matrix.py:376-401):
    reshape to (n, prod(rest) * ncols) and run one GEMM / diagonal scaling over the first axis."""
    nrest = int(np.prod(rest_dims))
    m_in = matrix.shape[0] if adjoint else matrix.shape[1]
    m_out = matrix.shape[1] if adjoint else matrix.shape[0]
    res_shape = (m_out * nrest,) if v.ndim == 1 else (m_out * nrest,) + tuple(v.shape[1:])
    v2 = v.reshape(m_in, -1)
    if isinstance(matrix, (DiagonalMatrix, DenseMatrix, ScaledIdentityMatrix, PartialDiagonalMatrix)):
        out = matrix._apply_adjoint(v2) if adjoint else matrix._apply(v2)
        return out.reshape(res_shape)
    raise RuntimeError(f"Unsupported type{type(matrix)}!")


# ------------------------------------------------------------------------------------------------
def identity(n, dtype=np.float64) -> ScaledIdentityMatrix:
    """Create an identity matrix (matrix.py:404-408)."""
    n = int(n)
    return ScaledIdentityMatrix(n, dtype(1.0))


def matrix_hash(a) -> int:
    """Hash of a matrix (matrix.py:411-418)."""
    if isinstance(a, torch.Tensor):
        a = D.to_host(a)
    if isinstance(a, np.ndarray):
        return hash(np.ascontiguousarray(a).data.tobytes())
    elif np.isscalar(a):
        return hash(a)
    return a.hash()


def asmatrixtype(a) -> MatrixBase:
    assert isinstance(a, MatrixBase) or (_is_array(a) and a.ndim == 2)
    if _is_array(a):
        return DenseMatrix(a)
    return a


def _vecprod(v1: np.ndarray, v2: np.ndarray, size: Optional[int] = None) -> np.ndarray:
    """Elementwise product of two vectors, right-padded with zeros to ``size`` (matrix.py:429-439)."""
    assert isinstance(v1, np.ndarray)
    assert isinstance(v2, np.ndarray)
    n = min(v1.size, v2.size)
    return D.to_host(D.diag_mul(D.as_dev(v1[:n]), D.as_dev(v2[:n]), size if size is not None else n))


def _pad_by_zero(arr: np.ndarray, size: int) -> np.ndarray:
    assert arr.size <= size
    if arr.size == size:
        return arr
    res = np.zeros(size, dtype=arr.dtype)
    res[0:arr.size] = arr
    return res


# ---- addition: name-based double dispatch like matrix.py:453-513 -------------------------------
def _dense_sum(a: MatrixBase, b: MatrixBase) -> DenseMatrix:
    return DenseMatrix(D.axpby(1.0, a._dense_dev(), 1.0, b._dense_dev()))


def _add_DiagonalMatrix_DenseMatrix(a, b):
    return _dense_sum(b, a)


def _add_DiagonalMatrix_DiagonalMatrix(a, b):
    return DiagonalMatrix(D.axpby(1.0, a._diag_dev(), 1.0, b._diag_dev()))


def _add_DiagonalMatrix_PartialDiagonalMatrix(a, b):
    a_diag = a.diagonals.reshape(b.matrix.shape[0], -1)
    if np.allclose(a_diag, a_diag[:, 0:1]):
        return PartialDiagonalMatrix(b.matrix + DiagonalMatrix(np.ascontiguousarray(a_diag[:, 0])), b.rest_dims)
    return _dense_sum(a, b)


def _add_PartialDiagonalMatrix_PartialDiagonalMatrix(a, b):
    if a.rest_dims == b.rest_dims:
        return PartialDiagonalMatrix(a.matrix + b.matrix, a.rest_dims)
    return _dense_sum(a, b)


def _add_DenseMatrix_DenseMatrix(a, b):
    return _dense_sum(a, b)


def _add_ScaledIdentityMatrix_ScaledIdentityMatrix(a, b):
    return ScaledIdentityMatrix(a.shape, a.coeff + b.coeff)


def _add_ScaledIdentityMatrix_DiagonalMatrix(a, b):
    return DiagonalMatrix(D.axpby(1.0, a._diag_dev(), 1.0, b._diag_dev()))


def _add_ScaledIdentityMatrix_PartialDiagonalMatrix(a, b):
    return PartialDiagonalMatrix(ScaledIdentityMatrix(b.matrix.shape[0], a.coeff) + b.matrix, b.rest_dims)


def _add_matrices(a, b):
    assert isinstance(a, MatrixBase)
    assert isinstance(b, MatrixBase)
    assert tuple(a.shape) == tuple(b.shape)
    f1 = globals().get(f"_add_{type(a).__name__}_{type(b).__name__}")
    if f1 is not None:
        return f1(a, b)
    f2 = globals().get(f"_add_{type(b).__name__}_{type(a).__name__}")
    if f2 is not None:
        return f2(b, a)
    return _dense_sum(a, b)
