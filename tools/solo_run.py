"""Cluster-resident solve of cfg2 (timing / ncu): python tools/solo_run.py [niter] [rtol]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admmsolver_b200 import batch, problems  # noqa: E402

niter = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
p = problems.spm_single(problems.ir_basis(), Nw=2000)
e = batch.SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True)
for n in (niter, niter, 20 * niter, 20 * niter):
    e.reset(mu=p.mu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    e.solve(n, rtol=0.0)
    e1.record()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it = int(e.iters[0])
    print("solve(%d): wall %.3f ms, device %.3f ms, %.3f us/iteration (device), iters %d" %
          (n, dt * 1e3, e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / max(it, 1), it))
