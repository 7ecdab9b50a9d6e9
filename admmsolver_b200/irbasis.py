"""On-device intermediate-representation (IR) basis and the SpM pipeline either side of the solver
(spm.ipynb:52-65 basis, :155-163 expansion of the spectrum, :198-199 sampling matrix,
:214-219 sum rule, :300 reconstruction).

The reference takes the basis from ``sparse_ir`` (absent here).  ``problems.ir_basis`` restates it on the host
with ``np.linalg.svd``; this module does the same construction with the arithmetic on the GPU:

* the singular value decomposition of the quadrature-weighted kernel by ``admm_svd_jacobi`` -- one-sided Jacobi on a
  cooperative grid, accurate to a few ulps *relative to every singular value*, which is what a kernel whose
  singular values span 16 decades needs (LAPACK's divide-and-conquer is accurate relative to the largest one);
* every projection -- ``v_l(omega)`` on an arbitrary grid, ``G(tau) -> g_l``, ``rho(omega) -> rho_l``, ``rho(omega) =
  v(omega) . x0`` -- as a tensor-core GEMM (``admm_gemm``).

Quadrature nodes, weights and the kernel samples themselves are input construction (host NumPy, a few MB uploaded
once).  PyTorch is used for device memory only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _dev as D
from ._lib import OP_N
from . import problems

__all__ = ["DeviceIRBasis", "ir_basis_device"]


@dataclass
class DeviceIRBasis:
    beta: float
    wmax: float
    s: torch.Tensor          # (L,) singular values, descending
    tau: np.ndarray          # quadrature nodes / weights (host: they only parametrise kernel samples)
    wtau: np.ndarray
    u_tau: torch.Tensor      # (L, ntau)   u_l(tau_i)
    omega: np.ndarray
    womega: np.ndarray
    v_omega: torch.Tensor    # (L, nomega) v_l(omega_j)
    sweeps: int = 0

    @property
    def size(self) -> int:
        return int(self.s.numel())

    def _uw(self) -> torch.Tensor:
        return (self.u_tau * torch.from_numpy(self.wtau).to(self.u_tau.device)[None, :]).contiguous()

    def v(self, omega: np.ndarray) -> torch.Tensor:
        """v_l(omega) on an arbitrary grid, (L, n) on the device: (1/s_l) int u_l(tau) K(tau, omega) dtau."""
        K = D.as_dev(problems._kernel(self.tau, np.asarray(omega, dtype=float), self.beta))
        return (D.gemm(OP_N, self._uw(), K) / self.s[:, None]).contiguous()

    def sampling_matrix(self, omega: np.ndarray, symmetric: bool = False) -> torch.Tensor:
        """P (n, L) with P[j, l] = v_l(omega_j): the coupling matrix of the non-negativity block (spm.ipynb:198-199).

        ``symmetric``: on a grid with omega[n-1-j] = -omega[j] the basis functions have the parity of their index,
        v_l(-w) = (-1)^l v_l(w), numerically to ~1e-9; project P onto EXACT parity (P[n-1-j, l] = (-1)^l P[j, l] bit for
        bit), which lets the fused SpM engine fold its pass over pairs of sampling points (half the tensor work)."""
        P = self.v(omega).t().contiguous()
        if symmetric:
            sign = torch.ones(P.shape[1], dtype=P.dtype, device=P.device)
            sign[1::2] = -1.0
            Ps = 0.5 * (P + P.flip(0) * sign[None, :])
            if float((Ps - P).abs().max()) > 1e-6 * float(P.abs().max()):
                raise ValueError("the basis functions do not have the parity of their index on this grid")
            P = Ps.contiguous()
        return P

    def sum_rule(self) -> torch.Tensor:
        """C (1, L), C_l = int v_l(omega) domega (spm.ipynb:214-219)."""
        w = torch.from_numpy(self.womega).to(self.v_omega.device)[:, None].contiguous()
        return D.gemm(OP_N, self.v_omega, w).t().contiguous()

    def expand_spectrum(self, rho_on_nodes) -> torch.Tensor:
        """rho_l = int v_l(omega) rho(omega) domega for spectra sampled on the basis' own omega nodes:
        (nomega,) or (nomega, nb) -> (L,) or (L, nb) (spm.ipynb:155-163)."""
        r = D.as_dev(rho_on_nodes)
        one = r.ndim == 1
        r2 = (r[:, None] if one else r) * torch.from_numpy(self.womega).to(r.device)[:, None]
        out = D.gemm(OP_N, self.v_omega, r2.contiguous())
        return out[:, 0].contiguous() if one else out

    def project_gtau(self, g_on_nodes) -> torch.Tensor:
        """g_l = int u_l(tau) G(tau) dtau for Green's functions sampled on the basis' own tau nodes: (ntau,) or
        (ntau, nb), real or complex -> (L,) or (L, nb): the data vector of the least-squares term."""
        g = D.as_dev(g_on_nodes)
        one = g.ndim == 1
        g2 = (g[:, None] if one else g).contiguous()
        out = D.gemm(OP_N, self._uw(), g2)
        return out[:, 0].contiguous() if one else out

    def reconstruct(self, x0, omega: np.ndarray) -> torch.Tensor:
        """rho(omega_j) = sum_l v_l(omega_j) x0_l (spm.ipynb:300): (L,) or (L, nb) -> (n,) or (n, nb)."""
        x = D.as_dev(x0)
        one = x.ndim == 1
        x2 = (x[:, None] if one else x).contiguous()
        P = self.sampling_matrix(omega)
        out = D.gemm(OP_N, P, x2)
        return out[:, 0].contiguous() if one else out

    def to_host(self) -> "problems.IRBasis":
        return problems.IRBasis(self.beta, self.wmax, self.s.cpu().numpy(), self.tau, self.wtau, self.u_tau.cpu().numpy(),
                                self.omega, self.womega, self.v_omega.cpu().numpy())


def ir_basis_device(beta: float = 100.0, wmax: float = 10.0, eps: float = 1e-7, max_sweeps: int = 48) -> DeviceIRBasis:
    """Same construction as ``problems.ir_basis`` (composite 16-point Gauss-Legendre panels in tau and omega, SVD of
    sqrt(w_tau) K sqrt(w_omega), keep s_l / s_0 > eps, sign convention: the largest-magnitude sample of every v_l is
    positive) with the decomposition and the scalings on the device."""
    tau, wtau, omega, womega = problems.ir_quadrature(beta, wmax)
    K = problems._kernel(tau, omega, beta)
    Kw = D.as_dev(np.sqrt(wtau)[:, None] * K * np.sqrt(womega)[None, :])
    U, s, V, sweeps = D.svd_jacobi(Kw, max_sweeps=max_sweeps, return_sweeps=True)
    L = int((s / s[0] > eps).sum().item())
    s = s[:L].contiguous()
    dev = Kw.device
    u_tau = (U[:, :L] / torch.from_numpy(np.sqrt(wtau)).to(dev)[:, None]).t().contiguous()
    v_omega = (V[:, :L] / torch.from_numpy(np.sqrt(womega)).to(dev)[:, None]).t().contiguous()
    j = v_omega.abs().argmax(dim=1)
    sign = torch.sign(v_omega[torch.arange(L, device=dev), j])
    return DeviceIRBasis(beta, wmax, s, tau, wtau, (u_tau * sign[:, None]).contiguous(), omega, womega,
                         (v_omega * sign[:, None]).contiguous(), sweeps)
