"""CPU oracle: flat NumPy restatement of the reference ADMM loop.

TEST INFRASTRUCTURE ONLY.  Nothing in ``admmsolver_b200`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path is CUDA-only.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (a) the reference's known-answer vector ``notebooks/basis_pursuit.ipynb:
137-138`` and (b) golden outputs of the unmodified reference generated in the
build container by ``tests/golden/make_golden.py`` (committed ``.npz`` files).

The reference evaluates the loop through generic operator objects
(``optimizer.py:302-341``); specialised to the two coupling patterns used by
BASELINE.json's configurations it reduces to the closed forms below.  A
Cholesky solve replaces ``np.linalg.inv`` (``matrix.py:77-78``), which changes
rounding at the 1e-14 level only.

Pattern A -- ``[LeastSquares(alpha, A, y), L1Regularizer(lam, N)]`` with the
condition ``(1, 0, I, I)`` (``test_optimizer.py:74-79``).

Pattern B -- ``[ConstrainedLeastSquares(alpha, A0, g, C, D), L1Regularizer(lam, L),
NonNegativePenalty(Nw)]`` with ``(0, 1, I, I)``, ``(0, 2, P, I)``
(``spm.ipynb:243-259``); batches are packed batch-fastest like
``PartialDiagonalMatrix`` (``matrix.py:301-401``), i.e. arrays of shape (L, nb).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
from scipy.linalg import cho_factor, cho_solve

__all__ = ["soft_threshold", "project_plus", "bp_solve", "spm_solve", "spm_solve_independent", "psd_project",
           "BPState", "SpMState"]


def soft_threshold(y: np.ndarray, lam) -> np.ndarray:
    """``_softmax`` (objectivefunc.py:335-355): strict comparisons, zero otherwise."""
    out = np.zeros_like(y)
    up = y > lam
    dn = y < -lam
    out[up] = (y - lam)[up]
    out[dn] = (y + lam)[dn]
    return out


def project_plus(x: np.ndarray) -> np.ndarray:
    """``_project_plus`` (objectivefunc.py:330-333)."""
    out = x.copy()
    out[x < 0] = 0
    return out


def _nrm(a: np.ndarray) -> float:
    return float(np.linalg.norm(np.ravel(a)))


def _rel_lt(diff: float, a: float, b: float, rtol: float) -> bool:
    """``norm(d)/max(norm(p1), norm(p2)) < rtol`` (optimizer.py:244-247); 0/0 -> nan -> False."""
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        return bool(np.float64(diff) / np.float64(max(a, b)) < rtol)


def _mu_step(mu: float, primal: float, dual: float, max_mu: float, fact: float = 2.0,
             th: float = 10.0) -> float:
    """``update_mu`` for one pair (optimizer.py:295-299)."""
    if primal > th * dual:
        mu *= fact
    if dual > th * primal:
        mu /= fact
    return min(mu, max_mu)


# --------------------------------------------------------------------------
# Pattern A: basis pursuit / LASSO
# --------------------------------------------------------------------------
@dataclass
class BPState:
    x0: np.ndarray
    x1: np.ndarray
    h: np.ndarray
    mu: float
    primal: List[float] = field(default_factory=list)
    dual: List[float] = field(default_factory=list)
    mu_hist: List[float] = field(default_factory=list)
    niter_done: int = 0
    converged: bool = False

    def objective(self, A, y, alpha, lam) -> float:
        """``SimpleOptimizer.__call__`` (optimizer.py:171-173) on the current x."""
        return float(alpha * np.linalg.norm(y - A @ self.x0) ** 2 + lam * np.sum(np.abs(self.x1)))


def bp_solve(A: np.ndarray, y: np.ndarray, alpha: float, lam: float, niter: int,
             mu: float = 1.0, max_mu: float = 1e3, interval_update_mu: int = 100,
             rtol: float = 1e-12, update_h: bool = True,
             state: Optional[BPState] = None) -> BPState:
    """Pattern A through ``SimpleOptimizer.solve`` (optimizer.py:302-320).

    x-update: ``LeastSquares.solve`` (objectivefunc.py:98-110) with
    ``h = -h10 - mu x1`` (``_hk``, optimizer.py:194-200); z-update:
    ``L1Regularizer.solve`` (objectivefunc.py:174-195) with ``h = h10 - mu x0``
    (optimizer.py:183-189); dual ascent optimizer.py:334-341.
    """
    M, N = A.shape
    cplx = np.iscomplexobj(A) or np.iscomplexobj(y)
    Ah = A.conj().T
    AhA = Ah @ A
    Aty = alpha * (Ah @ y)
    st = state or BPState(np.zeros(N, complex), np.zeros(N, complex), np.zeros(N, complex), mu)
    fac_mu = None
    fac = None
    for it in range(niter):
        x0_old = st.x0.copy()
        if fac_mu != st.mu:
            fac = cho_factor(alpha * AhA + st.mu * np.eye(N))
            fac_mu = st.mu
        st.x0 = cho_solve(fac, Aty + st.h + st.mu * st.x1).astype(complex)
        t = st.x0.real - st.h.real / st.mu
        st.x1 = soft_threshold(t, np.full(N, 0.5 * lam / st.mu)).astype(complex)
        if update_h:
            st.h = st.h + st.mu * (st.x1 - st.x0)
        n_pd = _nrm(st.x0 - st.x1)
        n_dd = _nrm(st.mu * (st.x0 - x0_old))
        st.primal.append(n_pd)
        st.dual.append(n_dd)
        st.niter_done += 1
        conv = _rel_lt(n_pd, _nrm(st.x0), _nrm(st.x1), rtol) and \
            _rel_lt(n_dd, _nrm(st.mu * st.x0), _nrm(st.mu * x0_old), rtol)
        if conv:
            st.converged = True
            st.mu_hist.append(st.mu)
            return st
        if it % interval_update_mu == 0:
            st.mu = _mu_step(st.mu, n_pd, n_dd, max_mu)
        st.mu_hist.append(st.mu)
    return st


# --------------------------------------------------------------------------
# Pattern B: SpM (single problem or packed batch with batch-wide norms)
# --------------------------------------------------------------------------
@dataclass
class SpMState:
    x0: np.ndarray
    x1: np.ndarray
    x2: np.ndarray
    h10: np.ndarray
    h20: np.ndarray
    mu10: float
    mu20: float
    primal: List[float] = field(default_factory=list)
    dual: List[float] = field(default_factory=list)
    mu_hist: List[tuple] = field(default_factory=list)
    niter_done: int = 0
    converged: bool = False

    def objective(self, s, g, lam, alpha=1.0) -> float:
        """alpha ||g - (-diag(s)) x0||^2 + lam |x1|_1 (+ 0 for the non-negative term)."""
        sv = s if self.x0.ndim == 1 else s[:, None]
        return float(alpha * np.linalg.norm(np.ravel(g + sv * self.x0)) ** 2
                     + lam * np.sum(np.abs(self.x1)))


def spm_solve(s: np.ndarray, P: np.ndarray, C: np.ndarray, D: np.ndarray, g: np.ndarray,
              lam: float, niter: int, mu: float = 0.1, alpha: float = 1.0, max_mu: float = 1e3,
              interval_update_mu: int = 100, rtol: float = 1e-12, update_h: bool = True,
              state: Optional[SpMState] = None, allreduce=None) -> SpMState:
    """Pattern B.  ``g`` is (L,) for one problem or (L, nb) for a packed batch.

    In the packed case every norm runs over the whole packed vector, so mu and
    the stopping test are batch-global -- exactly what the reference does when
    the operators are ``PartialDiagonalMatrix``.

    x0: ``ConstrainedLeastSquares.solve`` (objectivefunc.py:138-157) with
    ``A = -diag(s)``; x1: L1 prox; x2: ``NonNegativePenalty.solve``
    (objectivefunc.py:256-271); residuals optimizer.py:251-274; convergence
    optimizer.py:232-249; mu update optimizer.py:277-299.

    ``allreduce``: optional callable summing a float64 vector over ranks.  When the batch is
    sharded, every rank passes its slab of ``g`` and the ten squared-norm partials are summed
    across ranks before the square roots -- the multi-GPU batch-wide criterion.
    """
    L = s.size
    Nw = P.shape[0]
    single = g.ndim == 1
    G_ = (g[:, None] if single else g).astype(complex)
    nb = G_.shape[1]
    nc = C.shape[0]
    assert C.shape == (nc, L)
    if nc == 1:
        Dv = np.broadcast_to(np.asarray(D, dtype=float).ravel(), (nb,)) if np.size(D) in (1, nb) else None
        assert Dv is not None
    else:
        # several constraint rows (objectivefunc.py:148-157 in general form): D is (nc,) or (nc, nb)
        Dv = np.asarray(D, dtype=float)
        Dv = np.broadcast_to(Dv.reshape(nc, -1), (nc, nb))
    c = C[0]
    PtP = P.T @ P
    b0 = -alpha * s[:, None] * G_                       # alpha * A^H y with A = -diag(s)
    if state is None:
        st = SpMState(np.zeros((L, nb), complex), np.zeros((L, nb), complex),
                      np.zeros((Nw, nb), complex), np.zeros((L, nb), complex),
                      np.zeros((Nw, nb), complex), mu, mu)
    else:
        st = state
        for name in ("x0", "x1", "x2", "h10", "h20"):
            a = getattr(st, name)
            if a.ndim == 1:
                setattr(st, name, a[:, None].copy())
    key = None
    fac = w = sigma = None
    for it in range(niter):
        x0_old = st.x0.copy()
        if key != (st.mu10, st.mu20):
            Gm = alpha * np.diag(s * s) + st.mu10 * np.eye(L) + st.mu20 * PtP
            fac = cho_factor(Gm)
            w = cho_solve(fac, c.conj())
            sigma = c @ w
            if nc > 1:
                W = cho_solve(fac, C.conj().T)              # (L, nc)
                Sinv = np.linalg.inv(C @ W)
            key = (st.mu10, st.mu20)
        rhs = b0 + st.h10 + st.mu10 * st.x1 + P.T @ (st.h20 + st.mu20 * st.x2)
        xi1 = cho_solve(fac, rhs)
        if nc == 1:
            st.x0 = xi1 + w[:, None] * ((Dv - c @ xi1) / sigma)[None, :]
        else:
            st.x0 = xi1 + W @ (Sinv @ (Dv - C @ xi1))
        t1 = st.x0.real - st.h10.real / st.mu10
        st.x1 = np.where(t1 > 0.5 * lam / st.mu10, t1 - 0.5 * lam / st.mu10,
                         np.where(t1 < -0.5 * lam / st.mu10, t1 + 0.5 * lam / st.mu10, 0.0)
                         ).astype(complex)
        Px0 = P @ st.x0
        t2 = Px0.real - st.h20.real / st.mu20
        st.x2 = np.where(t2 < 0, 0.0, t2).astype(complex)
        if update_h:
            st.h10 = st.h10 + st.mu10 * (st.x1 - st.x0)
            st.h20 = st.h20 + st.mu20 * (st.x2 - Px0)
        Px0_old = P @ x0_old
        if allreduce is None:
            p10 = _nrm(st.x0 - st.x1)
            d10 = _nrm(st.mu10 * (st.x0 - x0_old))
            p20 = _nrm(Px0 - st.x2)
            d20 = _nrm(st.mu20 * (Px0 - Px0_old))
            nx0, nx1, nxo = _nrm(st.x0), _nrm(st.x1), _nrm(x0_old)
            nPx0, nx2, nPxo = _nrm(Px0), _nrm(st.x2), _nrm(Px0_old)
        else:
            sq = np.array([_nrm(v) ** 2 for v in (st.x0 - st.x1, st.x0 - x0_old, Px0 - st.x2, Px0 - Px0_old,
                                                  st.x0, st.x1, x0_old, Px0, st.x2, Px0_old)])
            sq = np.sqrt(allreduce(sq))
            p10, d10, p20, d20 = sq[0], st.mu10 * sq[1], sq[2], st.mu20 * sq[3]
            nx0, nx1, nxo, nPx0, nx2, nPxo = sq[4:]
        st.primal.append(p10 + p20)
        st.dual.append(d10 + d20)
        st.niter_done += 1
        conv = _rel_lt(p10, nx0, nx1, rtol) and \
            _rel_lt(d10, st.mu10 * nx0, st.mu10 * nxo, rtol) and \
            _rel_lt(p20, nPx0, nx2, rtol) and \
            _rel_lt(d20, st.mu20 * nPx0, st.mu20 * nPxo, rtol)
        if conv:
            st.converged = True
            st.mu_hist.append((st.mu10, st.mu20))
            break
        if it % interval_update_mu == 0:
            st.mu10 = _mu_step(st.mu10, p10, d10, max_mu)
            st.mu20 = _mu_step(st.mu20, p20, d20, max_mu)
        st.mu_hist.append((st.mu10, st.mu20))
    if single:
        for name in ("x0", "x1", "x2", "h10", "h20"):
            setattr(st, name, getattr(st, name)[:, 0])
    return st


def spm_solve_independent(s, P, C, D, g, lam, niter, **kw) -> List[SpMState]:
    """Per-problem mode: every column of ``g`` is its own reference instance."""
    nb = g.shape[1]
    Dv = np.broadcast_to(np.asarray(D, dtype=float).ravel(), (nb,))
    return [spm_solve(s, P, C, np.array([Dv[b]]), g[:, b], lam, niter, **kw) for b in range(nb)]


def psd_project(h: np.ndarray, diagonals: np.ndarray, shape, axis: int) -> np.ndarray:
    """``SemiPositiveDefinitePenalty.solve`` (objectivefunc.py:312-327): x = -Re(h)/mu reshaped to the
    3-way ``shape``; every slice along ``axis`` is replaced by its projection onto the PSD cone, the
    symmetric matrix being defined by the LOWER triangle of the slice (``np.linalg.eigh`` default)."""
    x = (-(np.real(h) / diagonals)).reshape(shape)
    x = np.moveaxis(x, axis, 0).copy()
    for i in range(x.shape[0]):
        lo = np.tril(x[i])
        sym = lo + np.tril(x[i], -1).T
        ev, U = np.linalg.eigh(sym)
        keep = ev >= 0
        x[i] = (U[:, keep] * ev[keep]) @ U[:, keep].T
    return np.moveaxis(x, 0, axis).ravel()
