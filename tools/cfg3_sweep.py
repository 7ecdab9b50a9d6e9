"""Launch-configuration sweep for small SpM batches: time per iteration for the balanced decomposition
(nsplit=None) with 1 or 2 problem tiles per warp, and for classic row splits."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from admmsolver_b200 import batch, problems
basis = problems.ir_basis()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = problems.spm_batch(min(nb, 4096), basis, Nw=2000, seed=1000)
g = np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb]
for mt, nsplit in ((1, None), (2, None), (1, 3), (2, 4)):
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, mt=mt, nsplit=nsplit)
    e.solve(100); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5): e.solve(100)
    t1.record(); torch.cuda.synchronize()
    print(f"nb={nb} mt={mt} nsplit={nsplit} -> dims nsplit={e.dims.nsplit} nbal={e.dims.nbal}: {t0.elapsed_time(t1)/500*1e3:.1f} us/iter  {nb*500/(t0.elapsed_time(t1)*1e-3)/1e6:.1f} M problem-iters/s")
