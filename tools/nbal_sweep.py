"""cfg3: time per iteration of the balanced decomposition vs number of pieces (CTAs): python tools/nbal_sweep.py [nb]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems

basis = problems.ir_basis()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = problems.spm_batch(min(nb, 4096), basis, Nw=2000, seed=1000)
g = np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb]
for mt, nbal in ((1, None), (1, 444), (1, 518), (1, 592), (1, 740), (2, 444)):
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, mt=mt, nbal=nbal)
    e.solve(100)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        e.solve(100)
    t1.record()
    torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / 500 * 1e3
    print(f"nb={nb} mt={mt} nbal={nbal} -> dims mt={e.dims.mt} nsplit={e.dims.nsplit} nbal={e.dims.nbal}: {us:.1f} us/iter  {nb / us:.1f} M problem-iters/s")
