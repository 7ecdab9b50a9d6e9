"""Helpers with the API of the reference's ``admmsolver.util`` (/root/reference/src/admmsolver/util.py).

``norm`` is the only one on the solve path (optimizer.py:10); the two finite-difference builders
construct *inputs* (smoothness regularisers) on the host and never run inside the loop.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _dev as D


def norm(x) -> float:
    """2-norm of a vector (util.py:40); NumPy arrays and CUDA tensors are reduced on the device."""
    return D.norm(D.as_dev(x))


def second_deriv_prj(x: np.ndarray) -> np.ndarray:
    """(N-2) x N matrix P with y''(x_i) ~= sum_j P_ij y_j on a non-uniform increasing grid
    (three-point stencil; util.py:4-23)."""
    x = np.asarray(x, dtype=np.float64)
    assert np.all(np.diff(x) > 0), "x must be in increasing order!"
    n = x.size
    fwd = x[2:] - x[1:-1]
    bwd = x[1:-1] - x[:-2]
    c = 2.0 / (fwd * bwd * (fwd + bwd))
    prj = np.zeros((n - 2, n))
    rows = np.arange(n - 2)
    prj[rows, rows] = c * fwd
    prj[rows, rows + 1] = -c * (fwd + bwd)
    prj[rows, rows + 2] = c * bwd
    return prj


def smooth_regularizer_coeff(omega: np.ndarray) -> np.ndarray:
    r"""Matrix R with \int |y''|^2 d\omega ~= |R y|^2 (util.py:26-39)."""
    omega = np.asarray(omega, dtype=np.float64)
    assert np.all(np.diff(omega) > 0), "omega must be in increasing order!"
    dx = 0.5 * (omega[2:] - omega[:-2])
    return np.sqrt(dx)[:, None] * second_deriv_prj(omega)
