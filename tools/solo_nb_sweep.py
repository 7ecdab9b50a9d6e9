"""Per-problem criterion: cluster-resident solve (one 8-CTA cluster per problem, several waves) vs the batch kernels
as a function of the batch size: python tools/solo_nb_sweep.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems

basis = problems.ir_basis()
p = problems.spm_batch(512, basis, Nw=2000, seed=1000)
batch.SharedSpM.SOLO_MAX_NB = 100000
for nb in (16, 32, 64, 128, 256, 512):
    g = p.g[:, :nb]
    for solo in (True, False):
        e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=False)
        e.solve(200, use_solo=solo)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            e.reset(mu=p.mu)
            e.solve(500, use_solo=solo)
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / 1500 * 1e3
        print(f"nb={nb} solo={solo}: {us:.1f} us per iteration of the batch, {nb / us:.2f} M problem-iters/s")
