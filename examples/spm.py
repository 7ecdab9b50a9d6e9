#!/usr/bin/env python
"""notebooks/spm.ipynb of SpM-lab/admmsolver (sparse-modelling analytic continuation) through the
drop-in API, with the plotting removed and `sparse_ir` replaced by the self-contained IR-like basis of
`admmsolver_b200.irbasis.ir_basis_device` (SVD of the fermionic kernel on Gauss-Legendre panels by one-sided Jacobi
on the GPU; L = 39 for beta = 100, wmax = 10, eps = 1e-7 as at spm.ipynb:214).  SURVEY.md 8(f) row f4: the pipeline
either side of the solver -- spectral function -> IR coefficients rho_l -> g_l = -s_l rho_l (+ noise) before,
rho(omega) = v(omega) . x0 after -- runs on the device (`admm_svd_jacobi`, `admm_gemm`); `device_basis=False` takes
the host construction (`problems.ir_basis`, np.linalg.svd) instead.

    PYTHONPATH=.:compat python examples/spm.py [niter]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "compat")]

import numpy as np  # noqa: E402

from admmsolver.matrix import DenseMatrix, DiagonalMatrix, identity  # noqa: E402
from admmsolver.objectivefunc import ConstrainedLeastSquares, L1Regularizer, NonNegativePenalty  # noqa: E402
from admmsolver.optimizer import Problem, SimpleOptimizer  # noqa: E402
from admmsolver_b200 import irbasis, problems  # noqa: E402


def main(niter: int = 10000, nw: int = 2000, verbose: bool = True, device_basis: bool = True):
    # spm.ipynb:61-65  basis
    wmax, beta = 10.0, 100.0
    rho = problems.rho_three_gaussians
    smpl_w = np.linspace(-wmax, wmax, nw)
    if device_basis:
        dbasis = irbasis.ir_basis_device(beta=beta, wmax=wmax, eps=1e-7)
        basis = dbasis.to_host()
        # spm.ipynb:155-163  IR expansion of the model spectrum; :198-199 sampling matrix; :214-219 sum rule -- device GEMMs
        rhol = dbasis.expand_spectrum(rho(dbasis.omega)).cpu().numpy()
        prj_w = dbasis.sampling_matrix(smpl_w).cpu().numpy()
        prj_sum = dbasis.sum_rule().cpu().numpy()
    else:
        basis = problems.ir_basis(beta=beta, wmax=wmax, eps=1e-7)
        # spm.ipynb:104-107,155-163  model spectrum and its IR expansion on the basis' own quadrature
        rhol = basis.v_omega @ (basis.womega * rho(basis.omega))
        # spm.ipynb:198-199  sampling points in real frequency (uniform grid instead of the roots of v_{L-1})
        prj_w = np.ascontiguousarray(basis.v(smpl_w).T)
        # spm.ipynb:214-219  sum rule
        prj_sum = basis.sum_rule()
    L = basis.size
    gl = -basis.s * rhol
    # spm.ipynb:243-259  the problem
    alpha, noise = 1e-4, 1e-4
    gl_dirty = gl + noise * np.random.RandomState(0).randn(L)
    lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(basis.s), gl_dirty, prj_sum, np.array([1]))
    l1 = L1Regularizer(alpha, L)
    nn = NonNegativePenalty(prj_w.shape[0])
    problem = Problem([lstsq, l1, nn], [(0, 1, identity(L), identity(L)), (0, 2, prj_w, identity(prj_w.shape[0]))])
    # spm.ipynb:281-285
    opt = SimpleOptimizer(problem, mu=0.1)
    opt.solve(niter)
    x0 = opt.x[0]
    # spm.ipynb:300  reconstruction rho(omega) = v(omega) . x0 on a fine grid, on the device
    omegas = np.linspace(-5, 5, 1000)
    if device_basis:
        rho_rec = dbasis.reconstruct(x0, omegas).cpu().numpy().real
    else:
        rho_rec = np.asarray((DenseMatrix(np.ascontiguousarray(basis.v(omegas).T)) @ x0)).real
    err = np.abs(rho_rec - rho(omegas)).max()
    if verbose:
        print("L =", L, " Nw =", nw, " iterations run =", len(opt._primal_residual))
        print("sum rule  prj_sum @ x0 =", (prj_sum @ x0)[0])                     # spm.ipynb:270 -> (1+0j)
        print("max |rho_rec - rho|    =", err, " (max rho = %.3f)" % rho(omegas).max())
        print("min rho_rec            =", rho_rec.min())
        print("residuals (last)       = primal %.3e, dual %.3e" % (opt._primal_residual[-1], opt._dual_residual[-1]))
    return dict(opt=opt, rho_rec=rho_rec, rho=rho(omegas), sum_rule=(prj_sum @ x0)[0], L=L)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
