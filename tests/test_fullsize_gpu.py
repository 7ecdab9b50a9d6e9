"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 2^20
problems): the full batch is R replicas of a small batch the oracle can solve, so

  * every replica of a problem must come out bit-identical wherever it sits in the batch (layout, tiling,
    CTA scheduling and the TMA ring cannot depend on the position);
  * per-problem mode: a replica equals the independent oracle solve of that problem (1e-10);
  * batch-wide mode: all ten squared norms scale by R, so the ratios that drive mu and the stopping test are
    those of the small packed batch -- the full batch must follow the oracle's packed solve (1e-10, same mu);
  * the sum rule C x0 = D holds for every problem (enforced exactly by the KKT step).
"""
import numpy as np
import pytest
import torch

from conftest import rel

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("symmetric", [False, True])
def test_spm_sweep_full_size(build_lib, ir_basis, symmetric):
    """cfg5 size: 2^20 complex SpM problems, L = 39, Nw = 2000 (16.9 GB of state), both criteria; with the sampling
    matrix as the quadrature delivers it (plain pass) and projected onto the exact parity of the IR basis (folded pass)."""
    from admmsolver_b200 import batch, problems
    from oracle import flat
    nb, nd, niter, interval = 1 << 20, 64, 60, 25
    p = problems.spm_batch(nd, ir_basis, Nw=2000, seed=77, symmetric=symmetric)
    g_dev = torch.from_numpy(p.g).cuda().repeat(1, nb // nd).contiguous()       # replica r of problem d at column r*nd + d
    # ---- batch-wide criterion (the packed reference semantics)
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g_dev, lam=p.lam, mu=p.mu, batch_wide=True)
    assert e.dims.nsplit == 1 and e.dims.nbal == 0 and e.dims.mt == 2          # the fused step kernel
    assert e.fold == symmetric
    e.solve(niter, interval_update_mu=interval)
    x0 = e.x0_device()
    st = flat.spm_solve(p.s, p.P, p.C, np.ones(nd), p.g, p.lam, niter, mu=p.mu, interval_update_mu=interval)
    assert float(e.mu10[0]) == st.mu10 and float(e.mu20[0]) == st.mu20
    first = x0[:, :nd]
    assert rel(first.cpu().numpy(), st.x0) < TOL
    assert bool((x0.view(39, nb // nd, nd) == first[:, None, :]).all())          # replicas bit-identical
    C = torch.from_numpy(p.C[0]).cuda().to(x0.dtype)
    assert float(((C[None, :] @ x0)[0] - 1.0).abs().max()) < 1e-12              # sum rule, all 2^20 problems
    assert rel(np.sqrt(nb // nd) * np.asarray(st.primal), np.asarray(e.primal_residual)) < 1e-8
    del e, x0
    torch.cuda.empty_cache()
    # ---- per-problem criterion (independent reference instances), every problem on its own mu history
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g_dev, lam=p.lam, mu=p.mu, batch_wide=False)
    e.solve(niter, interval_update_mu=interval)
    x0 = e.x0_device()
    first = x0[:, :nd]
    assert bool((x0.view(39, nb // nd, nd) == first[:, None, :]).all())
    mu20 = e.mu20[:nb].view(nb // nd, nd)
    assert bool((mu20 == mu20[:1]).all())
    seen = set()
    for d in (0, 7, 8, 33, 63):
        sb = flat.spm_solve(p.s, p.P, p.C, np.array([1.0]), p.g[:, d], p.lam, niter, mu=p.mu, interval_update_mu=interval)
        assert rel(first[:, d].cpu().numpy(), sb.x0) < TOL
        assert float(e.mu10[d]) == sb.mu10 and float(e.mu20[d]) == sb.mu20
        seen.add((sb.mu10, sb.mu20))
    assert len(seen) >= 1


def test_bp_cfg4_full_size(build_lib):
    """cfg4 size: 65 536 basis-pursuit problems, each with its own 128 x 512 A (34 GB of A)."""
    from admmsolver_b200 import batch, problems
    from oracle import flat
    nb, nd, niter = 65536, 128, 120
    A, y, _ = problems.basis_pursuit_batch(nd, 128, 512, 10, seed0=500)
    A_dev = torch.from_numpy(A).cuda().repeat(nb // nd, 1, 1).contiguous()
    y_dev = torch.from_numpy(y).cuda().repeat(nb // nd, 1).contiguous()
    e = batch.BatchedBasisPursuit(A_dev, y_dev, 1.0, 0.1)
    assert e.At is not None                                                    # the single-sweep kernel
    e.solve(niter)
    x0 = e._x0
    first = x0[:nd]
    assert bool((x0.view(nb // nd, nd, 512) == first[None]).all())             # replicas bit-identical
    assert bool((e.mu.view(nb // nd, nd) == e.mu[:nd][None]).all())
    assert bool((e.iters == niter).all())
    for d in (0, 1, 64, 127):
        st = flat.bp_solve(A[d], y[d], 1.0, 0.1, niter)
        assert rel(first[d].cpu().numpy(), st.x0.real) < TOL
        assert float(e.mu[d]) == st.mu
