#!/usr/bin/env python
"""notebooks/basis_pursuit.ipynb of SpM-lab/admmsolver, cell by cell, through the drop-in API
(`import admmsolver` resolves to compat/admmsolver -> admmsolver_b200; the whole loop runs on the GPU).

    PYTHONPATH=.:compat python examples/basis_pursuit.py

Prints the two known-answer numbers of the notebook (basis_pursuit.ipynb:137-138):
max|xanswer| = 1.4312955709975443, max|xanswer - x0| = 0.0054070107628...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "compat")]

import numpy as np  # noqa: E402

from admmsolver.matrix import identity  # noqa: E402
from admmsolver.objectivefunc import L1Regularizer, LeastSquares  # noqa: E402
from admmsolver.optimizer import Problem, SimpleOptimizer  # noqa: E402


def main(niter: int = 100, verbose: bool = True):
    # basis_pursuit.ipynb:63-69
    np.random.seed(1234)
    N, M, K = 1000, 100, 20
    A = np.random.randn(M, N)
    xanswer = np.zeros(N)
    xanswer[:K] = np.random.randn(K)
    xanswer = np.random.permutation(xanswer)
    y = A @ xanswer
    # basis_pursuit.ipynb:85-101
    alpha = 0.1
    lstsq = LeastSquares(1.0, A, y)
    l1 = L1Regularizer(alpha, N)
    problem = Problem([lstsq, l1], [(1, 0, identity(N), identity(N))])
    opt = SimpleOptimizer(problem)
    opt.solve(niter)
    x0 = opt.x[0]
    if verbose:
        print("max |xanswer|      =", np.abs(xanswer).max())
        print("max |xanswer - x0| =", np.abs(xanswer - x0).max())
        print("objective          =", opt(opt.x))
    return xanswer, x0, opt


if __name__ == "__main__":
    main()
