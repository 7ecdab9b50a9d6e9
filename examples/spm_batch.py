#!/usr/bin/env python
"""A whole SpM analytic-continuation batch on the GPU, end to end (the workflow of notebooks/spm.ipynb for many
Green's functions at once -- k-points x orbitals -- sharing one IR basis; nothing of size Nw x nb ever visits the host):

    IR basis (Jacobi SVD of the kernel)  ->  G(tau) of nb model spectra  ->  g_l = int u_l(tau) G(tau) dtau
    ->  ADMM solve of all problems (fused shared-A engine, per-problem penalties and stopping)
    ->  rho(omega) = v(omega) . x0 on a fine grid

    PYTHONPATH=. python examples/spm_batch.py [nb] [niter]

With `moments=True` the first moment of every spectrum is imposed next to the sum rule (a two-row constraint C x0 = D).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from admmsolver_b200 import _dev as Dv  # noqa: E402
from admmsolver_b200 import irbasis, problems  # noqa: E402
from admmsolver_b200._lib import OP_N  # noqa: E402
from admmsolver_b200.batch import SharedSpM  # noqa: E402


def main(nb: int = 256, niter: int = 2000, nw: int = 1000, noise: float = 1e-5, moments: bool = False, verbose: bool = True):
    wmax, beta = 10.0, 100.0
    basis = irbasis.ir_basis_device(beta=beta, wmax=wmax, eps=1e-7)            # on the device
    L = basis.size
    dev = basis.s.device
    # model spectra: random three-Gaussian mixtures on the basis' own omega quadrature (nomega x nb)
    rs = np.random.RandomState(7)
    cen, wid = rs.uniform(-2.0, 2.0, (nb, 3)), rs.uniform(0.4, 1.0, (nb, 3))
    wgt = rs.dirichlet(np.ones(3), nb)
    wq = basis.omega
    rho = np.zeros((wq.size, nb))
    for k in range(3):
        rho += wgt[None, :, k] * np.exp(-((wq[:, None] - cen[None, :, k]) / wid[None, :, k]) ** 2) / (np.sqrt(np.pi) * wid[None, :, k])
    # G(tau_i) = -int K(tau_i, omega) rho(omega) domega  on the tau quadrature, then noise
    K = torch.from_numpy(problems._kernel(basis.tau, basis.omega, beta) * basis.womega[None, :]).to(dev)
    G = -Dv.gemm(OP_N, K.contiguous(), torch.from_numpy(rho).to(dev))          # input construction (ntau x nb)
    G = G + noise * torch.from_numpy(rs.randn(*G.shape)).to(dev)
    g_l = basis.project_gtau(G)                                                # (L, nb): tensor-core GEMM
    # operators of the problem
    omega = np.linspace(-wmax, wmax, nw)
    P = basis.sampling_matrix(omega, symmetric=True)                           # (nw, L) on the device, exact parity -> folded pass
    C = basis.sum_rule()                                                       # (1, L)
    D = torch.ones(1, nb, dtype=torch.float64, device=dev)
    if moments:
        # first moment of every spectrum: int omega rho(omega) domega = -(dG/dtau jump ...) -- here taken from the model
        wv = torch.from_numpy(basis.womega * basis.omega).to(dev)[:, None]
        C = torch.cat([C, Dv.gemm(OP_N, basis.v_omega, wv).t()], dim=0)        # (2, L)
        m1 = torch.from_numpy((basis.womega * basis.omega) @ rho).to(dev)[None, :]
        D = torch.cat([D, m1], dim=0)
    eng = SharedSpM(basis.s, P, C, D, g_l, lam=1e-5, mu=0.1, batch_wide=False)
    eng.solve(niter)
    x0 = eng.x0_device()                                                       # (L, nb) complex128 on the device
    rho_rec = basis.reconstruct(x0.real.contiguous(), omega)                   # (nw, nb) on the device
    rho_true = np.zeros((nw, nb))
    for k in range(3):
        rho_true += wgt[None, :, k] * np.exp(-((omega[:, None] - cen[None, :, k]) / wid[None, :, k]) ** 2) / (np.sqrt(np.pi) * wid[None, :, k])
    err = (rho_rec - torch.from_numpy(rho_true).to(dev)).abs().amax(dim=0) / torch.from_numpy(rho_true.max(axis=0)).to(dev)
    viol = (Dv.gemm(OP_N, C.contiguous(), x0.real.contiguous()) - D).abs().max()
    # how well the data are reproduced: |g_l + s_l x0_l| / |g_l| per spectrum (the least-squares term of the model)
    fit = ((g_l + basis.s[:, None] * x0.real).norm(dim=0) / g_l.norm(dim=0)).max()
    out = dict(L=L, nb=nb, max_rel_err=float(err.max()), median_rel_err=float(err.median()), min_rho=float(rho_rec.min()),
               constraint_violation=float(viol), data_misfit=float(fit), iters=eng.iters[:nb].cpu().numpy(), folded=eng.fold)
    if verbose:
        print(f"L = {L}, {nb} spectra, {nw} sampling points, iterations {out['iters'].min()}..{out['iters'].max()}")
        print(f"max |rho_rec - rho| / max rho: median {out['median_rel_err']:.3f}, worst {out['max_rel_err']:.3f}")
        print(f"min rho_rec = {out['min_rho']:.2e}, max |C x0 - D| = {out['constraint_violation']:.2e}, "
              f"worst data misfit |g + s x0| / |g| = {out['data_misfit']:.2e}")
    return out


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 2000)
