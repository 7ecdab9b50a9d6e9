"""One cluster-resident solve of cfg2 (for ncu): python tools/solo_run.py [niter]"""
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
for _ in range(3):
    e.reset(mu=p.mu)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.solve(niter)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("solve(%d): %.3f ms, %.2f us/iteration, iters %d" % (niter, dt * 1e3, dt * 1e6 / niter, int(e.iters[0])))
