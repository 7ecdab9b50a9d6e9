"""Whole-column fused step vs fused balanced step at the per-GPU batch sizes of the 4- and 8-GPU sweep."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from admmsolver_b200 import batch, problems
basis = problems.ir_basis()
p = problems.spm_batch(4096, basis, Nw=2000, seed=1000)
for nb in [int(a) for a in sys.argv[1:]] or [131072, 262144]:
    g = torch.from_numpy(np.tile(p.g, (1, -(-nb // 4096)))[:, :nb].copy()).cuda()
    for kw in (dict(), dict(mt=2, nbal=444), dict(mt=2, nbal=296), dict(mt=1, nbal=444)):
        e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, **kw)
        e.solve(30); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3): e.solve(100)
        t1.record(); torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / 300 * 1e3
        print(f"nb={nb} {kw} -> mt={e.dims.mt} nsplit={e.dims.nsplit} nbal={e.dims.nbal} step_mode={e._step_mode}: {us:.1f} us/iter "
              f"{nb / us:.1f} M problem-iters/s frac {nb * 324158.0 / (us * 1e-6) / 35.4e12:.3f}", flush=True)
        del e
