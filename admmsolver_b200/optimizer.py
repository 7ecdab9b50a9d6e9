"""ADMM driver with the API of the reference's ``admmsolver.optimizer``
(/root/reference/src/admmsolver/optimizer.py): ``EqualityCondition``, ``Model``/``Problem`` and
``SimpleOptimizer`` with ``solve / one_sweep / residual / check_convergence / update_mu``, the
``x`` property, ``__call__`` and the de-facto public ``_primal_residual``, ``_dual_residual``,
``_mu``, ``_h``, ``_x_old``.

Execution:

* ``solve()`` recognises the two coupling patterns of BASELINE.json and hands the whole loop to a
  fused CUDA engine (``batch.BatchedBasisPursuit`` -- pattern A; ``batch.SharedSpM`` -- pattern B,
  also when the operators are ``PartialDiagonalMatrix``-packed batches);
* every other model (and the step-wise public methods) runs on the generic device executor below:
  the reference's sweep restated over device vectors, with adjoints, ``E @ x`` products and
  ``mu_k`` matrices cached instead of rebuilt on every call.

Either way the state lives on the GPU during the loop and is mirrored into the NumPy arrays of
``opt.x`` / ``opt._h`` / ``opt._mu`` when control returns to Python.
"""
from __future__ import annotations

from itertools import product
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _dev as D
from .matrix import (DenseMatrix, DiagonalMatrix, MatrixBase, PartialDiagonalMatrix, ScaledIdentityMatrix,
                     asmatrixtype)
from .objectivefunc import (ConstrainedLeastSquares, L1Regularizer, LeastSquares, NonNegativePenalty,
                            ObjectiveFunctionBase)

__all__ = ["EqualityCondition", "Model", "Problem", "SimpleOptimizer"]


class EqualityCondition(object):
    """E1 @ x_{i1} - E2 @ x_{i2} = 0 with i1 != i2 (optimizer.py:12-38)."""

    def __init__(self, i1: int, i2: int, E1, E2) -> None:
        assert i1 != i2, "i1 != i2!"
        assert E1.shape[0] == E2.shape[0], "Leading dimensions of E1 and E2 do not match!"
        assert E1.ndim == 2
        assert E2.ndim == 2
        super().__init__()
        self.i1 = i1
        self.i2 = i2
        self.E1 = asmatrixtype(E1)
        self.E2 = asmatrixtype(E2)

    @property
    def size(self) -> int:
        return self.E1.shape[0]


class Model(object):
    """Sum of terms plus pairwise equality conditions (optimizer.py:40-118).  ``E[a, b]`` is the
    matrix that multiplies ``x_b`` in its condition with partner ``a``."""

    def __init__(self, functions: Sequence[ObjectiveFunctionBase],
                 equality_conditons: Union[tuple, List[EqualityCondition]] = ()) -> None:
        self._functions = functions
        self._num_func = len(functions)
        n = self._num_func
        self._E = np.full((n, n), None)
        self._EcE = np.full((n, n), None)
        self._EcE2 = np.full((n, n), None)
        for ie, e in enumerate(equality_conditons):
            try:
                self._add_equality_condition(EqualityCondition(*e) if isinstance(e, tuple) else e)
            except Exception:
                print(f"Error occured when adding {ie}-th equality condition!")
                raise
        for i in range(n):
            for k in range(n):
                if self._E[k, i] is None:
                    continue
                self._EcE[k, i] = self._E[i, k].T.conjugate() @ self._E[k, i]
                self._EcE2[k, i] = self._E[k, i].T.conjugate() @ self._E[k, i]

    @property
    def functions(self) -> Sequence[ObjectiveFunctionBase]:
        return self._functions

    @property
    def num_func(self) -> int:
        return self._num_func

    @property
    def E(self) -> np.ndarray:
        return self._E

    @property
    def EcE(self) -> np.ndarray:
        """E[i,k]^dagger E[k,i] at (k, i)"""
        return self._EcE

    @property
    def EcE2(self) -> np.ndarray:
        """E[k,i]^dagger E[k,i] at (k, i)"""
        return self._EcE2

    def _add_equality_condition(self, e: EqualityCondition) -> None:
        assert isinstance(e, EqualityCondition)
        assert e.E1.shape[1] == self._functions[e.i1].size_x, f"{e.E1.shape} {self._functions[e.i1].size_x}"
        assert e.E2.shape[1] == self._functions[e.i2].size_x, f"{e.E2.shape} {self._functions[e.i2].size_x}"
        if self._E[e.i1, e.i2] is not None:
            raise RuntimeError("Duplicate entries in equality_conditions")
        self._E[e.i2, e.i1] = e.E1
        self._E[e.i1, e.i2] = e.E2


Problem = Model   # backward compatibility (optimizer.py:117-118)


def _sum(objs):
    assert isinstance(objs, list)
    res = objs[0]
    for x in objs[1:]:
        res = res + x if isinstance(res, MatrixBase) else D.axpby(1.0, res, 1.0, x)
    return res


def _is_identity(m) -> bool:
    return isinstance(m, ScaledIdentityMatrix) and m.is_diagonal() and complex(m.coeff) == 1.0


def _real_dense(m: MatrixBase) -> Optional[torch.Tensor]:
    """Device tensor of a real Dense/Diagonal/ScaledIdentity operator, else None."""
    if isinstance(m, (DenseMatrix, DiagonalMatrix, ScaledIdentityMatrix)):
        t = m._dense_dev()
        return None if t.is_complex() else t
    return None


# ------------------------------------------------------------------------------------------------
# fused-plan recognition
# ------------------------------------------------------------------------------------------------
class _BPPlan:
    """Pattern A: [LeastSquares(real dense A, real y), L1Regularizer], condition (1,0,I,I) or (0,1,I,I)."""

    def __init__(self, model: Model):
        from .batch import BatchedBasisPursuit
        ls, l1 = model.functions
        A = ls._A._dense_dev()
        self.eng = BatchedBasisPursuit(A, ls._y_dev, alpha=float(ls._alpha), lam=float(l1._alpha), keep_history=True,
                                       keep_x_old=True)

    @staticmethod
    def match(model: Model) -> bool:
        f = model.functions
        if model.num_func != 2 or type(f[0]) is not LeastSquares or type(f[1]) is not L1Regularizer:
            return False
        if not isinstance(f[0]._A, DenseMatrix) or f[0]._A._dense_dev().is_complex() or f[0]._y_dev.is_complex():
            return False
        m, n = f[0]._A.shape
        if D._lib.lib.admm_bp_supported(int(m), int(n)) == 0:      # the N-vectors have to fit one CTA's shared memory
            return False
        return _is_identity(model.E[0, 1]) and _is_identity(model.E[1, 0])


class _SpMPlan:
    """Pattern B: [ConstrainedLeastSquares, L1Regularizer, NonNegativePenalty] with (0,1,I,I) and
    (0,2,P,I); single problem or PartialDiagonalMatrix-packed batch (batch-wide mu / stopping)."""

    def __init__(self, model: Model, mu10: float, mu20: float, max_mu: float):
        from .batch import SharedSpM
        cls, l1, nn = model.functions
        info = self.match(model)
        nb = info["nb"]
        unwrap = (lambda m: m.matrix) if info["packed"] else (lambda m: m)
        A = unwrap(cls._A)
        L = A.shape[1]
        G0 = D.gemm(2, A._dense_dev(), A._dense_dev())                   # A^H A (L x L, real)
        G0 = D.axpby(float(cls._alpha), G0)
        b0 = cls._Aty.reshape(L, nb)
        P = unwrap(model.E[2, 0])._dense_dev()
        Cm = unwrap(cls._C)._dense_dev()                                  # (constraint rows, L): up to four rows
        self.nb, self.L, self.Nw = nb, L, P.shape[0]
        self.eng = SharedSpM.from_operators(G0, b0, P, Cm, cls._D_dev, lam=float(l1._alpha), mu10=mu10, mu20=mu20,
                                            batch_wide=True, max_mu=max_mu, force_complex=True, keep_x_old=True)

    @staticmethod
    def match(model: Model):
        f = model.functions
        if model.num_func != 3 or type(f[0]) is not ConstrainedLeastSquares or type(f[1]) is not L1Regularizer \
                or type(f[2]) is not NonNegativePenalty:
            return None
        E = model.E
        if E[1, 2] is not None or E[2, 1] is not None or E[1, 0] is None or E[2, 0] is None:
            return None
        if not (_is_identity(E[1, 0]) and _is_identity(E[0, 1]) and _is_identity(E[0, 2])):
            return None
        cls = f[0]
        P, A, Cm = E[2, 0], cls._A, cls._C
        packed = isinstance(P, PartialDiagonalMatrix)
        if packed:
            if not (isinstance(A, PartialDiagonalMatrix) and isinstance(Cm, PartialDiagonalMatrix)):
                return None
            # any number of batch axes (k-points x orbitals ...): the packed vector is C-ordered, so the
            # flattened rest index is the batch index
            if not (len(P.rest_dims) >= 1 and tuple(P.rest_dims) == tuple(A.rest_dims) == tuple(Cm.rest_dims)):
                return None
            nb = int(np.prod(P.rest_dims))
            P, A, Cm = P.matrix, A.matrix, Cm.matrix
        else:
            nb = 1
        for m in (P, A, Cm):
            if _real_dense(m) is None:
                return None
        L = A.shape[1]
        if L > 64 or not (1 <= Cm.shape[0] <= 4) or cls._D_dev.numel() != Cm.shape[0] * nb:
            return None
        return {"nb": nb, "packed": packed}


# ------------------------------------------------------------------------------------------------
class SimpleOptimizer(object):
    """The simplest ADMM solver (optimizer.py:121-341)."""

    def __init__(self, model: Model, x0=None, mu=None, max_mu: float = 1e+3) -> None:
        assert isinstance(model, Model)
        num_func = model.num_func
        self._h = np.full((num_func, num_func), None)
        self._mu = np.full((num_func, num_func), 0.0)
        self._model = model
        self._max_mu = max_mu
        if x0 is not None:
            for i in range(len(x0)):
                assert model._functions[i].size_x == x0[i].size
            self._x = [np.array(x_, dtype=np.complex128) for x_ in x0]
        else:
            self._x = [np.zeros(model.functions[k].size_x, dtype=np.complex128) for k in range(num_func)]
        if mu is None:
            mu = 1.0
        self._pairs: List[Tuple[int, int]] = []
        for i, j in product(range(num_func), repeat=2):
            if model.E[i, j] is None or i <= j:
                continue
            self._h[i, j] = np.zeros(model.E[i, j].shape[0], dtype=np.complex128)
            self._mu[i, j] = mu
            self._pairs.append((i, j))
        self._primal_residual: List[float] = []
        self._dual_residual: List[float] = []

        # device state of the generic executor
        self._xd: Optional[List[torch.Tensor]] = None
        self._hd: Dict[Tuple[int, int], torch.Tensor] = {}
        self._xd_old: Optional[List[torch.Tensor]] = None
        self._snapshot = None
        self._Eadj: Dict[Tuple[int, int], MatrixBase] = {}
        self._muk_cache: Dict[int, Tuple[tuple, MatrixBase]] = {}
        # fused plan
        self._plan = None
        self._plan_kind = None
        if _BPPlan.match(model):
            self._plan_kind = "bp"
        elif _SpMPlan.match(model) is not None:
            self._plan_kind = "spm"

    # ------------------------------------------------------------------ public state
    @property
    def x(self) -> List[np.ndarray]:
        return self._x

    def __call__(self, x: List[np.ndarray]) -> float:
        """Evaluate the cost function (optimizer.py:171-173)."""
        return float(np.sum([f(x_) for x_, f in zip(x, self._model.functions)]))

    # ------------------------------------------------------------------ host <-> device mirrors
    def _host_state(self):
        return ([a.copy() for a in self._x], {p: self._h[p].copy() for p in self._pairs}, self._mu.copy())

    def _host_dirty(self) -> bool:
        if self._snapshot is None:
            return True
        xs, hs, mus = self._snapshot
        if not np.array_equal(mus, self._mu):
            return True
        if any(not np.array_equal(a, b) for a, b in zip(xs, self._x)):
            return True
        return any(not np.array_equal(hs[p], self._h[p]) for p in self._pairs)

    def _upload(self) -> None:
        if self._xd is not None and not self._host_dirty():
            return
        self._xd = [D.as_dev(a, D.C128) for a in self._x]
        self._hd = {p: D.as_dev(self._h[p], D.C128) for p in self._pairs}
        self._snapshot = self._host_state()

    def _download(self) -> None:
        for k, t in enumerate(self._xd):
            self._x[k][:] = D.to_host(t)
        for p in self._pairs:
            self._h[p][:] = D.to_host(self._hd[p])
        if self._xd_old is not None:
            self._x_old = [D.to_host(t) for t in self._xd_old]
        self._snapshot = self._host_state()

    # ------------------------------------------------------------------ generic executor pieces
    def _adj(self, i: int, k: int) -> MatrixBase:
        """E[i,k]^H, built once (the reference rebuilds it on every call, optimizer.py:187,198)."""
        if (i, k) not in self._Eadj:
            self._Eadj[(i, k)] = self._model.E[i, k].T.conjugate()
        return self._Eadj[(i, k)]

    def _hk(self, k: int):
        """Linear term for the x_k update (optimizer.py:175-207); device tensor or None."""
        self._upload()
        return self._hk_dev(k)

    def _hk_dev(self, k: int) -> Optional[torch.Tensor]:
        EcE = self._model.EcE
        res = []
        for i in range(k):
            if self._h[k, i] is None:
                continue
            res.append(D.axpby(1.0, self._adj(i, k) @ self._hd[(k, i)], -float(self._mu[k, i]), EcE[k, i] @ self._xd[i]))
        for i in range(k + 1, self._model.num_func):
            if self._h[i, k] is None:
                continue
            res.append(D.axpby(-1.0, self._adj(i, k) @ self._hd[(i, k)], -float(self._mu[i, k]), EcE[k, i] @ self._xd[i]))
        for r in res:
            assert r.numel() == self._model.functions[k].size_x, f"{r.numel()} {self._model.functions[k].size_x}"
        return _sum(res) if res else None

    def _mu_k(self, k: int) -> Optional[MatrixBase]:
        """Quadratic term for the x_k update (optimizer.py:209-230), cached per mu values."""
        EcE2 = self._model.EcE2
        partners = [(k, i) for i in range(k) if self._h[k, i] is not None] + \
                   [(i, k) for i in range(k + 1, self._model.num_func) if self._h[i, k] is not None]
        if not partners:
            return None
        key = tuple(float(self._mu[p]) for p in partners)
        hit = self._muk_cache.get(k)
        if hit is not None and hit[0] == key:
            return hit[1]
        res = []
        for (a, b) in partners:
            other = b if a == k else a
            res.append(float(self._mu[a, b]) * EcE2[other, k])
        m = _sum(res)
        self._muk_cache[k] = (key, m)
        return m

    def _pair_vectors(self, i: int, j: int):
        E = self._model.E
        p1 = E[i, j] @ self._xd[j]
        p2 = E[j, i] @ self._xd[i]
        return p1, p2

    def _dual_vectors(self, i: int, j: int, p1: torch.Tensor, scaled: bool = True):
        """mu * E[j,i] @ (E[i,j] @ x_j) for the current and the previous x_j (optimizer.py:241-242:
        E[j,i], not its adjoint -- reproduced as written).  ``scaled=False`` leaves the
        factor mu to the caller (norms are homogeneous: two launches less per pair and iteration)."""
        E = self._model.E
        d1 = E[j, i] @ p1
        d2 = E[j, i] @ (E[i, j] @ self._xd_old[j])
        if scaled:
            mu = float(self._mu[i, j])
            d1, d2 = D.axpby(mu, d1), D.axpby(mu, d2)
        return d1, d2

    # ------------------------------------------------------------------ public step-wise API
    def one_sweep(self, update_h: bool) -> None:
        """Update all variables in a single sweep (optimizer.py:322-341)."""
        self._upload()
        self._sweep_dev(update_h)
        self._download()

    def _sweep_dev(self, update_h: bool) -> None:
        """One Gauss-Seidel sweep + dual ascent (optimizer.py:322-341) on the device state, updated IN
        PLACE: the state tensors keep their addresses, so a sweep can be captured in a CUDA graph."""
        model = self._model
        if self._xd_old is None or any(o.shape != t.shape for o, t in zip(self._xd_old, self._xd)):
            self._xd_old = [torch.empty_like(t) for t in self._xd]
        for o, t in zip(self._xd_old, self._xd):
            o.copy_(t)
        for k in range(model.num_func):
            f = model.functions[k]
            h, mu = self._hk_dev(k), self._mu_k(k)
            if hasattr(f, "_solve_complex"):
                xk = f._solve_complex(h, mu)
            else:
                xk = f.solve(h, mu)
                if isinstance(xk, np.ndarray):        # user-defined term working on NumPy
                    xk = D.as_dev(xk)
            self._xd[k].copy_(xk)                     # (casts real results to the complex128 state)
        self._pv = {}            # E1 x_i, E2 x_j of this sweep: the norms that follow need the same vectors
        if update_h:
            for (i, j) in self._pairs:
                p1, p2 = self._pair_vectors(i, j)
                self._pv[(i, j)] = (p1, p2)
                self._hd[(i, j)].copy_(D.axpby(1.0, self._hd[(i, j)], float(self._mu[i, j]), D.axpby(1.0, p2, -1.0, p1)))

    def _need_old(self) -> None:
        if self._xd_old is None:
            if hasattr(self, "_x_old"):
                self._xd_old = [D.as_dev(a, D.C128) for a in self._x_old]
            else:
                raise AttributeError("'SimpleOptimizer' object has no attribute '_x_old'")

    def check_convergence(self, rtol) -> bool:
        """optimizer.py:232-249; 0/0 -> NaN -> not converged."""
        self._upload()
        self._need_old()
        converged = True
        for (i, j) in self._pairs:
            p1, p2 = self._pair_vectors(i, j)
            d1, d2 = self._dual_vectors(i, j, p1)
            with np.errstate(all="ignore"):
                converged = converged and bool(np.float64(D.norm(p1, p2)) / np.float64(max(D.norm(p1), D.norm(p2))) < rtol)
                converged = converged and bool(np.float64(D.norm(d1, d2)) / np.float64(max(D.norm(d1), D.norm(d2))) < rtol)
        return converged

    def _pair_residuals(self, i: int, j: int) -> Tuple[float, float]:
        p1, p2 = self._pair_vectors(i, j)
        d1, d2 = self._dual_vectors(i, j, p1)
        return D.norm(p1, p2), D.norm(d1, d2)

    def residual(self) -> Tuple[float, float]:
        """Primal and dual residual (optimizer.py:251-274)."""
        self._upload()
        self._need_old()
        primal = dual = 0.0
        for (i, j) in self._pairs:
            p, d = self._pair_residuals(i, j)
            primal += p
            dual += d
        return primal, dual

    def update_mu(self, fact_incr: float = 2.0, th_change: float = 10.0) -> None:
        """optimizer.py:277-299."""
        self._upload()
        self._need_old()
        for (i, j) in self._pairs:
            primal, dual = self._pair_residuals(i, j)
            if primal > th_change * dual:
                self._mu[i, j] *= fact_incr
            if dual > th_change * primal:
                self._mu[i, j] /= fact_incr
            self._mu[i, j] = min(self._mu[i, j], self._max_mu)
        if self._snapshot is not None:
            self._snapshot = (self._snapshot[0], self._snapshot[1], self._mu.copy())

    # ------------------------------------------------------------------ solve
    def solve(self, niter: int = 10000, callback: Optional[Callable] = None, interval_update_mu: int = 100,
              update_h: bool = True, rtol: float = 1e-12) -> None:
        """optimizer.py:302-320: sweep -> residual append -> callback -> convergence -> mu update."""
        if self._plan_kind is not None and update_h and callback is None and niter > 0:
            try:
                if self._solve_fused(niter, interval_update_mu, rtol):
                    return
            except NotImplementedError:
                self._plan_kind = None        # state not representable by the fused engine
        self._solve_generic(niter, callback, interval_update_mu, update_h, rtol)

    def _launch_norms(self, buf: torch.Tensor) -> None:
        """The six norms per coupled pair that residual(), check_convergence() and update_mu() need
        (optimizer.py:232-299), each pair/dual vector formed ONCE, squared norms into ``buf`` (no host
        synchronisation).  Row per pair: |p1-p2|, |p1|, |p2|, |d1-d2|, |d1|, |d2| -- the last three WITHOUT the
        factor mu of the dual vectors."""
        pv = getattr(self, "_pv", None) or {}
        for n, (i, j) in enumerate(self._pairs):
            p1, p2 = pv[(i, j)] if (i, j) in pv else self._pair_vectors(i, j)
            d1, d2 = self._dual_vectors(i, j, p1, scaled=False)       # _solve_generic multiplies the norms by mu
            D.pair_norms_into(buf, 6 * n, p1, p2, d1, d2)
        self._pv = {}

    def _graphable(self) -> bool:
        """Every term runs on the device (a user-defined NumPy term cannot be captured in a CUDA graph)."""
        return all(hasattr(f, "_solve_complex") or type(f).__module__.startswith("admmsolver_b200")
                   for f in self._model.functions)

    def _solve_generic(self, niter, callback, interval_update_mu, update_h, rtol) -> None:
        """Same order of operations as the reference loop; without a callback the residuals, the stopping
        test and update_mu() of an iteration share one set of norms (one host synchronisation per iteration
        instead of one per norm; the reference recomputes every E @ x for each of the three)."""
        self._upload()
        if not self._pairs and callback is None:
            # no equality conditions: the terms are independent, check_convergence() is vacuously True after the
            # first sweep (optimizer.py:232-249,313) -- one sweep, residuals (0, 0)
            if niter > 0:
                self._sweep_dev(update_h)
                self._primal_residual.append(0.0)
                self._dual_residual.append(0.0)
            self._download()
            return
        # The captured graph of the current penalties survives across solve() calls as long as the device state keeps
        # its addresses (capture + instantiation costs 10 ms and sporadically 100s of ms -- more than a short solve).
        dev = self._xd[0].device
        nb6 = 6 * max(1, len(self._pairs))
        if getattr(self, "_gnbuf", None) is None or self._gnbuf.numel() != nb6 or self._gnbuf.device != dev:
            self._gnbuf = torch.zeros(nb6, dtype=D.F64, device=dev)
            self._ggraphs, self._gwarm, self._gsig = {}, set(), None
        nbuf = self._gnbuf
        sig = (tuple(t.data_ptr() for t in self._xd) + tuple(self._hd[p].data_ptr() for p in self._pairs)
               + tuple(t.data_ptr() for t in (self._xd_old or [])))
        if sig != self._gsig:
            self._ggraphs, self._gwarm, self._gsig = {}, set(), sig
        graphable = callback is None and self._graphable()
        graph = graph_key = None
        npair = len(self._pairs)
        hist = None          # device copy of the norms of every iteration of a chunk
        it = 0
        while it < niter:
            if callback is not None:
                self._sweep_dev(update_h)
                primal, dual = self.residual()
                self._primal_residual.append(primal)
                self._dual_residual.append(dual)
                self._download()
                callback()
                self._upload()
                if self.check_convergence(rtol):
                    break
                if it % interval_update_mu == 0:
                    self.update_mu()
                it += 1
                continue
            # sweep + norms: eagerly the first time for every set of penalties (fills the adjoint / mu_k /
            # inverse caches outside any capture), then captured once and replayed as ONE graph launch
            key = tuple(float(self._mu[p]) for p in self._pairs) + (bool(update_h),)
            snap = None
            if graph_key != key:
                # one graph at a time: a captured graph references the per-penalty caches of the terms (inverse, mu_k),
                # which may be released once the penalties move on -- a change of key drops it and warms up again
                if key not in self._ggraphs:
                    self._ggraphs, self._gwarm = {}, set()
                graph, graph_key = self._ggraphs.get(key), key
            if graph is not None:
                # Run ahead: all iterations up to and including the next update_mu iteration are replayed back to
                # back, their norms collected on the device and read with ONE synchronisation.  Should the stopping
                # test have fired in the middle of the chunk (it can only happen once per solve), the state is rolled
                # back to the start of the chunk and exactly the iterations that count are replayed again.
                nxt = it if it % interval_update_mu == 0 else (it // interval_update_mu + 1) * interval_update_mu
                n = max(1, min(niter - it, nxt - it + 1, self.GENERIC_CHUNK))
                if n > 1:
                    snap = ([t.clone() for t in self._xd], {p: t.clone() for p, t in self._hd.items()},
                            [t.clone() for t in self._xd_old])
                    if hist is None:
                        hist = torch.empty(self.GENERIC_CHUNK, nbuf.numel(), dtype=D.F64, device=nbuf.device)
                    for k in range(n):
                        graph.replay()
                        hist[k].copy_(nbuf)
                    rows = np.sqrt(hist[:n].cpu().numpy())
                else:
                    graph.replay()
                    rows = np.sqrt(nbuf.cpu().numpy())[None]
            else:
                n = 1
                self._sweep_dev(update_h)
                self._launch_norms(nbuf)
                if graphable and key in self._gwarm:
                    before = D._lib.launch_count
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, capture_error_mode="thread_local"):
                            self._sweep_dev(update_h)
                            self._launch_norms(nbuf)
                        graph = g
                        self._ggraphs = {key: g}
                    except Exception as exc:          # something in the model synchronises: stay eager
                        graphable = False
                        torch.cuda.synchronize()
                        import warnings
                        warnings.warn("generic executor: CUDA-graph capture failed (%s); running eagerly" % (exc,))
                    D._lib.launch_count = before
                self._gwarm.add(key)
                rows = np.sqrt(nbuf.cpu().numpy())[None]
            rows = rows.reshape(n, npair, 6).copy()
            for n_, pk in enumerate(self._pairs):                       # the dual norms were taken without the factor mu
                rows[:, n_, 3:6] *= float(self._mu[pk])
            rows = rows.reshape(n, npair * 6)
            done = 0
            conv = False
            for k in range(n):
                nr = rows[k].reshape(npair, 6)
                self._primal_residual.append(float(sum(float(v) for v in nr[:, 0])))
                self._dual_residual.append(float(sum(float(v) for v in nr[:, 3])))
                done = k + 1
                with np.errstate(all="ignore"):
                    conv = all(bool(r[0] / max(r[1], r[2]) < rtol) and bool(r[3] / max(r[4], r[5]) < rtol) for r in nr)
                if conv:
                    break
            if conv and done < n:
                for dst, src in zip(self._xd, snap[0]):
                    dst.copy_(src)
                for pk, src in snap[1].items():
                    self._hd[pk].copy_(src)
                for dst, src in zip(self._xd_old, snap[2]):
                    dst.copy_(src)
                for _ in range(done):
                    graph.replay()
            it += done
            if conv:
                break
            if (it - 1) % interval_update_mu == 0:
                nr = rows[done - 1].reshape(npair, 6)
                for n_, (i, j) in enumerate(self._pairs):
                    primal, dual = float(nr[n_, 0]), float(nr[n_, 3])
                    if primal > 10.0 * dual:
                        self._mu[i, j] *= 2.0
                    if dual > 10.0 * primal:
                        self._mu[i, j] /= 2.0
                    self._mu[i, j] = min(self._mu[i, j], self._max_mu)
                if self._snapshot is not None:
                    self._snapshot = (self._snapshot[0], self._snapshot[1], self._mu.copy())
        self._download()

    # ---- fused engines
    def _solve_fused(self, niter: int, interval: int, rtol: float) -> bool:
        dirty = self._plan is None or self._host_dirty()
        if self._plan_kind == "bp":
            if self._plan is None:
                self._plan = _BPPlan(self._model)
            eng = self._plan.eng
            eng.max_mu = float(self._max_mu)
            if dirty:
                if any(np.abs(a.imag).max(initial=0.0) != 0.0 for a in (self._x[0], self._x[1], self._h[1, 0])):
                    raise NotImplementedError("complex state on the real basis-pursuit engine")
                eng.set_state(x0=self._x[0].real, x1=self._x[1].real, h=self._h[1, 0].real, mu=float(self._mu[1, 0]))
            n0 = len(eng.primal_residual[0])
            eng.solve(niter, interval_update_mu=interval, rtol=rtol)
            self._x[0][:] = eng.x0()[0]
            self._x[1][:] = eng.x1()[0]
            self._h[1, 0][:] = eng.h()[0]
            self._mu[1, 0] = float(eng.mu[0].item())
            self._primal_residual.extend(eng.primal_residual[0][n0:])
            self._dual_residual.extend(eng.dual_residual[0][n0:])
            # `_x_old` (optimizer.py:324): the state before the last sweep.  Only x_old[0] enters residual() /
            # check_convergence() / update_mu() of this pattern (pair (1,0): E[0,1] @ E[1,0] @ x_old[0]); the engine
            # hands it out exactly.  x_old[1] is not kept by the kernels: the current x1 stands in.
            if len(eng.primal_residual[0]) > n0:
                self._x_old = [eng.x0_old()[0].astype(np.complex128), self._x[1].copy()]
        else:
            if self._plan is None:
                self._plan = _SpMPlan(self._model, float(self._mu[1, 0]), float(self._mu[2, 0]), float(self._max_mu))
                dirty = any(np.any(a != 0) for a in self._x) or any(np.any(self._h[p] != 0) for p in self._pairs)
            eng = self._plan.eng
            L, Nw, nb = self._plan.L, self._plan.Nw, self._plan.nb
            if dirty:
                eng.mu10[:nb] = float(self._mu[1, 0])
                eng.mu20[:nb] = float(self._mu[2, 0])
                eng._refresh_slots()
                eng.set_state(x0=self._x[0].reshape(L, nb), x1=self._x[1].reshape(L, nb), x2=self._x[2].reshape(Nw, nb),
                              h10=self._h[1, 0].reshape(L, nb), h20=self._h[2, 0].reshape(Nw, nb))
            n0 = len(eng.primal_residual)
            eng.solve(niter, interval_update_mu=interval, rtol=rtol)
            self._x[0][:] = eng.x0().ravel()
            self._x[1][:] = eng.x1().ravel()
            self._x[2][:] = eng.x2().ravel()
            self._h[1, 0][:] = eng.h10().ravel()
            self._h[2, 0][:] = eng.h20().ravel()
            self._mu[1, 0] = float(eng.mu10[0].item())
            self._mu[2, 0] = float(eng.mu20[0].item())
            self._primal_residual.extend(eng.primal_residual[n0:])
            self._dual_residual.extend(eng.dual_residual[n0:])
            # `_x_old`: x_old[0] exact from the engine (both pairs (1,0) and (2,0) only read x_old[0]); the current
            # x1 / x2 stand in for x_old[1] / x_old[2], which the kernels do not keep (x2 is Nw x nb)
            if len(eng.primal_residual) > n0:
                self._x_old = [eng.x0_old().ravel().astype(np.complex128), self._x[1].copy(), self._x[2].copy()]
        self._xd = None          # the generic mirrors are stale now
        self._xd_old = None
        self._snapshot = self._host_state()
        return True


#: iterations the generic executor runs ahead between two host synchronisations (see _solve_generic)
SimpleOptimizer.GENERIC_CHUNK = 64
