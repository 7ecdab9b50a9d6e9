"""Deterministic synthetic inputs for the five BASELINE.json configurations.

Host-side input construction only (NumPy): nothing here is on the solve path.
The recipes follow SURVEY.md section 8(d):

* basis pursuit      -- reference ``test/test_optimizer.py:56-80`` and
  ``notebooks/basis_pursuit.ipynb`` (Gaussian A, K-sparse x, y = A x).
* SpM                -- reference ``notebooks/spm.ipynb``; the notebook needs
  ``sparse_ir`` (absent, no network) so an IR-like basis is built from the
  SVD of the fermionic analytic-continuation kernel on Gauss-Legendre panels.

Both the parity tests and ``bench.py`` import these generators so that the CUDA
engine, the oracle and the reference see bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

__all__ = [
    "basis_pursuit_instance",
    "basis_pursuit_batch",
    "IRBasis",
    "ir_basis",
    "spm_single",
    "spm_batch",
    "rho_three_gaussians",
]


# --------------------------------------------------------------------------
# basis pursuit
# --------------------------------------------------------------------------
def basis_pursuit_instance(M: int = 100, N: int = 1000, K: int = 20, seed: int = 1234
                           ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The notebook/test instance (defaults) or BASELINE cfg1 (M=200, K=10, seed=0).

    Uses the legacy global-seed call order of ``test_optimizer.py:62-72`` so the
    default arguments reproduce the reference's own golden vector
    (``basis_pursuit.ipynb:137-138``).
    """
    rs = np.random.RandomState(seed)
    A = rs.randn(M, N)
    xanswer = np.zeros(N)
    xanswer[:K] = rs.randn(K)
    xanswer = rs.permutation(xanswer)
    y = A @ xanswer
    return A, y, xanswer


def basis_pursuit_batch(nb: int, M: int = 128, N: int = 512, K: int = 10, seed0: int = 0
                        ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """cfg4: ``nb`` independent problems, problem ``b`` seeded with ``seed0 + b``."""
    A = np.empty((nb, M, N))
    y = np.empty((nb, M))
    xa = np.empty((nb, N))
    for b in range(nb):
        A[b], y[b], xa[b] = basis_pursuit_instance(M, N, K, seed0 + b)
    return A, y, xa


# --------------------------------------------------------------------------
# IR-like basis for SpM
# --------------------------------------------------------------------------
def _panel_nodes(edges: np.ndarray, order: int = 16) -> Tuple[np.ndarray, np.ndarray]:
    xg, wg = np.polynomial.legendre.leggauss(order)
    a = edges[:-1, None]
    b = edges[1:, None]
    x = 0.5 * (b - a) * (xg[None, :] + 1.0) + a
    w = 0.5 * (b - a) * wg[None, :]
    return x.ravel(), w.ravel()


def _kernel(tau: np.ndarray, omega: np.ndarray, beta: float) -> np.ndarray:
    """Fermionic kernel K(tau, omega) = exp(-tau*omega) / (1 + exp(-beta*omega)), overflow safe."""
    tau = np.asarray(tau)[:, None]
    omega = np.asarray(omega)[None, :]
    pos = omega >= 0
    out = np.empty(np.broadcast(tau, omega).shape)
    wp = np.where(pos, omega, 0.0)
    wn = np.where(pos, 0.0, omega)
    kp = np.exp(-tau * wp) / (1.0 + np.exp(-beta * wp))
    kn = np.exp((beta - tau) * wn) / (1.0 + np.exp(beta * wn))
    out = np.where(pos, kp, kn)
    return out


@dataclass
class IRBasis:
    beta: float
    wmax: float
    s: np.ndarray          # (L,) singular values
    tau: np.ndarray        # quadrature nodes in tau
    wtau: np.ndarray
    u_tau: np.ndarray      # (L, ntau)   u_l(tau_i)
    omega: np.ndarray      # quadrature nodes in omega
    womega: np.ndarray
    v_omega: np.ndarray    # (L, nomega) v_l(omega_j)

    @property
    def size(self) -> int:
        return int(self.s.size)

    def v(self, omega: np.ndarray) -> np.ndarray:
        """v_l(omega) on an arbitrary grid: (1/s_l) int u_l(tau) K(tau, omega) dtau -> (L, n)."""
        K = _kernel(self.tau, omega, self.beta)
        return ((self.u_tau * self.wtau[None, :]) @ K) / self.s[:, None]

    def sum_rule(self) -> np.ndarray:
        """C (1, L) with C_l = int v_l(omega) domega (== s_l (u_l(0) + u_l(beta)) up to the cut-off)."""
        return (self.v_omega @ self.womega)[None, :]


def ir_quadrature(beta: float, wmax: float):
    """Composite 16-point Gauss-Legendre nodes / weights of the basis construction: tau panels cluster toward 0 and
    beta, omega panels geometrically toward 0.  Returns (tau, wtau, omega, womega)."""
    npan_t = 24
    th = np.linspace(0.0, np.pi, npan_t + 1)
    edges_t = 0.5 * beta * (1.0 - np.cos(th))
    # refine toward both ends geometrically (the kernel varies on scale 1/wmax there)
    fine = beta * 0.5 ** np.arange(40, 4, -1)
    edges_t = np.unique(np.concatenate([edges_t, fine, beta - fine]))
    tau, wtau = _panel_nodes(edges_t)

    npan_w = 40
    geo = wmax * 0.5 ** np.arange(npan_w, -1, -1, dtype=float)
    edges_w = np.concatenate([-geo[::-1], [0.0], geo])
    omega, womega = _panel_nodes(edges_w)
    return tau, wtau, omega, womega


def ir_basis(beta: float = 100.0, wmax: float = 10.0, eps: float = 1e-7) -> IRBasis:
    """SVD basis of the kernel on composite 16-point Gauss-Legendre panels.

    tau panels cluster (Chebyshev-like) toward 0 and beta; omega panels are
    geometric toward 0.  Keeps s_l / s_0 > eps (L = 39 for beta=100, wmax=10,
    eps=1e-7, the size printed at ``spm.ipynb:214``).  Signs are fixed so that
    the largest-magnitude sample of every v_l is positive (SVD sign ambiguity).
    """
    tau, wtau, omega, womega = ir_quadrature(beta, wmax)
    K = _kernel(tau, omega, beta)
    Kw = np.sqrt(wtau)[:, None] * K * np.sqrt(womega)[None, :]
    U, s, Vt = np.linalg.svd(Kw, full_matrices=False)
    L = int(np.sum(s / s[0] > eps))
    s = s[:L]
    u_tau = (U[:, :L] / np.sqrt(wtau)[:, None]).T
    v_omega = Vt[:L, :] / np.sqrt(womega)[None, :]
    for l in range(L):
        j = int(np.argmax(np.abs(v_omega[l])))
        if v_omega[l, j] < 0:
            v_omega[l] *= -1.0
            u_tau[l] *= -1.0
    return IRBasis(beta, wmax, s, tau, wtau, u_tau, omega, womega, v_omega)


def _gaussian(x, mu, sigma):
    return np.exp(-((x - mu) / sigma) ** 2) / (np.sqrt(np.pi) * sigma)


def rho_three_gaussians(omega: np.ndarray) -> np.ndarray:
    """The model spectrum of ``spm.ipynb:104-107``."""
    return (0.2 * _gaussian(omega, 0.0, 0.15) + 0.4 * _gaussian(omega, 1.0, 0.8)
            + 0.4 * _gaussian(omega, -1.0, 0.8))


@dataclass
class SpMProblem:
    """One SpM analytic-continuation instance (or a batch sharing the basis).

    minimise  alpha_ls * || g - (-diag(s)) x0 ||^2 + lam * |x1|_1   s.t.  C x0 = D,
              x0 = x1,   P x0 = x2 >= 0          (``spm.ipynb:243-259``)
    """
    s: np.ndarray      # (L,)
    P: np.ndarray      # (Nw, L)   v_l(omega_j)
    C: np.ndarray      # (1, L)
    D: np.ndarray      # (1,) or (nb,)
    g: np.ndarray      # (L,) float64 | (L, nb) complex128
    lam: float         # L1 weight (the notebook's ``alpha``)
    mu: float          # initial penalty
    omega: np.ndarray  # (Nw,)
    rho_l: np.ndarray  # exact expansion coefficients, (L,) or (L, nb)


def spm_single(basis: IRBasis | None = None, Nw: int = 2000, noise: float = 1e-4, seed: int = 0,
               lam: float = 1e-4, mu: float = 0.1) -> SpMProblem:
    """cfg2: single SpM problem, L = basis.size, Nw uniform omega points in [-wmax, wmax]."""
    basis = basis or ir_basis()
    omega = np.linspace(-basis.wmax, basis.wmax, Nw)
    P = np.ascontiguousarray(basis.v(omega).T)
    rho_l = basis.v_omega @ (basis.womega * rho_three_gaussians(basis.omega))
    rs = np.random.RandomState(seed)
    g = -basis.s * rho_l + noise * rs.randn(basis.size)
    return SpMProblem(basis.s.copy(), P, basis.sum_rule(), np.array([1.0]), g, lam, mu, omega, rho_l)


def symmetrize_sampling(P: np.ndarray) -> np.ndarray:
    """Project a sampling matrix on a symmetric grid onto exact parity: P[Nw-1-r, l] = (-1)^l P[r, l]."""
    sign = np.ones(P.shape[1])
    sign[1::2] = -1.0
    return np.ascontiguousarray(0.5 * (P + P[::-1] * sign[None, :]))


def spm_batch(nb: int, basis: IRBasis | None = None, Nw: int = 2000, noise: float = 1e-4,
              seed: int = 0, lam: float = 1e-4, mu: float = 0.1, complex_noise: bool = True,
              symmetric: bool = False) -> SpMProblem:
    """cfg3/cfg5: ``nb`` spectra sharing one basis; random 3-Gaussian mixtures.

    Centres U(-2,2), widths U(0.1,1), Dirichlet(1,1,1) weights.  ``g`` is
    (L, nb) complex128 (batch index fastest, the packing of
    ``PartialDiagonalMatrix``, reference ``matrix.py:313-325,389``) with a
    seeded imaginary noise part.

    ``symmetric``: the IR basis functions have the parity of their index, v_l(-w) = (-1)^l v_l(w); the numerically
    built basis reproduces that on the symmetric grid only to ~1e-9.  With ``symmetric=True`` the sampling matrix is
    projected onto exact parity, P[Nw-1-r, l] = (-1)^l P[r, l] bit for bit -- what an analytic basis would deliver, and
    what lets the fused engine fold the pass over pairs of sampling points (``admm_spm_dims.fold``).
    """
    basis = basis or ir_basis()
    omega = np.linspace(-basis.wmax, basis.wmax, Nw)
    P = np.ascontiguousarray(basis.v(omega).T)
    if symmetric:
        P = symmetrize_sampling(P)
    rs = np.random.RandomState(seed)
    cen = rs.uniform(-2.0, 2.0, size=(nb, 3))
    wid = rs.uniform(0.1, 1.0, size=(nb, 3))
    wgt = rs.dirichlet(np.ones(3), size=nb)
    wq = basis.omega
    rho = np.zeros((nb, wq.size))
    for k in range(3):
        rho += wgt[:, k, None] * _gaussian(wq[None, :], cen[:, k, None], wid[:, k, None])
    rho_l = (basis.v_omega * basis.womega[None, :]) @ rho.T          # (L, nb)
    g = -basis.s[:, None] * rho_l + noise * rs.randn(basis.size, nb)
    g = g.astype(np.complex128)
    if complex_noise:
        g = g + 1j * noise * rs.randn(basis.size, nb)
    return SpMProblem(basis.s.copy(), P, basis.sum_rule(), np.ones(nb), g, lam, mu, omega, rho_l)
