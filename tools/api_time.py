"""User-level timing of the drop-in API on the two single-problem configurations (cfg1, cfg2)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admmsolver_b200 import problems  # noqa: E402
from admmsolver_b200.matrix import DenseMatrix, DiagonalMatrix, identity  # noqa: E402
from admmsolver_b200.objectivefunc import ConstrainedLeastSquares, L1Regularizer, LeastSquares, NonNegativePenalty  # noqa: E402
from admmsolver_b200.optimizer import Model, SimpleOptimizer  # noqa: E402


def clock(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out


# cfg1: basis pursuit 200 x 1000 (test_optimizer.py:52-82 with BASELINE sizes)
A, y, xa = problems.basis_pursuit_instance(200, 1000, 10, 0)
for rep in range(3):
    t_build, opt = clock(lambda: SimpleOptimizer(Model([LeastSquares(1.0, A, y), L1Regularizer(0.1, 1000)],
                                                       [(1, 0, identity(1000), identity(1000))])))
    t_solve, _ = clock(lambda: opt.solve(1000))
    t_x, x = clock(lambda: opt.x[0])
    print("cfg1 rep %d: build %.2f ms, solve(1000) %.2f ms, x %.2f ms, err %.2e" % (rep, t_build, t_solve, t_x, np.abs(x.real - xa).max()))

# cfg2: SpM single problem (spm.ipynb)
p = problems.spm_single(problems.ir_basis(), Nw=2000)
L, Nw = p.P.shape[1], p.P.shape[0]
for rep in range(3):
    def build():
        lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), p.g, p.C, np.array([1]))
        l1 = L1Regularizer(p.lam, L)
        nn = NonNegativePenalty(Nw)
        return SimpleOptimizer(Model([lstsq, l1, nn], [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]),
                               mu=p.mu)
    t_build, opt = clock(build)
    t_solve, _ = clock(lambda: opt.solve(1000))
    t_x, x = clock(lambda: opt.x[0])
    print("cfg2 rep %d: build %.2f ms, solve(1000) %.2f ms, x %.2f ms" % (rep, t_build, t_solve, t_x))
