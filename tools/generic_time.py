import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/compat"]
import numpy as np, torch
from admmsolver.matrix import identity, DiagonalMatrix, PartialDiagonalMatrix
from admmsolver.objectivefunc import LeastSquares, L1Regularizer, NonNegativePenalty
from admmsolver.optimizer import Model, SimpleOptimizer
import os
if os.environ.get('GENERIC_CHUNK'):
    SimpleOptimizer.GENERIC_CHUNK = int(os.environ['GENERIC_CHUNK'])
rs = np.random.RandomState(5)
n0, n1, n2 = 600, 500, 400
Ag = rs.randn(800, n0); yg = rs.randn(800)
E1 = rs.randn(n1, n0); Pg = rs.randn(n2, n0)
opt = SimpleOptimizer(Model([LeastSquares(1.3, Ag, yg), L1Regularizer(0.2, n1), NonNegativePenalty(n2)],
                            [(0, 1, E1, identity(n1)), (0, 2, Pg, DiagonalMatrix(np.linspace(1.0, 2.0, n2)))]), mu=0.7)
opt.solve(20); torch.cuda.synchronize()
t = time.perf_counter(); opt.solve(200); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("generic 3-term dense: %.1f us/iter" % (dt / 200 * 1e6))
# packed LASSO batch sharing one A (PartialDiagonalMatrix): the generic shared-A x-update GEMM
N, M, nbt = 256, 128, 512
A = rs.randn(M, N); Y = rs.randn(M, nbt)
lst = LeastSquares(1.0, PartialDiagonalMatrix(A, (nbt,)), Y.ravel())
opt = SimpleOptimizer(Model([lst, L1Regularizer(0.1, N * nbt)], [(1, 0, identity(N * nbt), identity(N * nbt))]))
opt.solve(10); torch.cuda.synchronize()
t = time.perf_counter(); opt.solve(100); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("packed LASSO %dx%d x %d problems: %.1f us/iter" % (M, N, nbt, dt / 100 * 1e6))
