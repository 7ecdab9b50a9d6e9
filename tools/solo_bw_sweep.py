"""Batch-wide criterion, a few packed problems: cluster-resident solve (global all-reduce of the norms per iteration)
vs the batch kernels: python tools/solo_bw_sweep.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems

basis = problems.ir_basis()
p = problems.spm_batch(32, basis, Nw=2000, seed=1000)
for nb in (2, 4, 8, 12, 16, 18, 20):
    g = p.g[:, :nb]
    for solo in (True, False):
        e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True)
        try:
            e.solve(200, use_solo=solo)
        except NotImplementedError as ex:
            print(f"nb={nb} solo={solo}: not supported ({str(ex)[:60]}...)")
            continue
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            e.reset(mu=p.mu)
            e.solve(1000, use_solo=solo)
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / 3000 * 1e3
        print(f"nb={nb} solo={solo}: {us:.1f} us per iteration of the batch, {nb / us:.2f} M problem-iters/s")
