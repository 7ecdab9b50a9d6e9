"""Launch-configuration sweep for the cfg3-sized SpM batch (4096 problems): time per iteration."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from admmsolver_b200 import batch, problems
basis = problems.ir_basis()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = problems.spm_batch(nb, basis, Nw=2000, seed=1000)
for mt in (1, 2):
    for nsplit in (1, 2, 3, 4, 5, 6, 7):
        try:
            e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), p.g, lam=p.lam, mu=p.mu, batch_wide=True, mt=mt, nsplit=nsplit)
        except Exception as ex:
            print(mt, nsplit, "skip", ex); continue
        e.solve(100); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): e.solve(100)
        t1.record(); torch.cuda.synchronize()
        print(f"mt={mt} nsplit={nsplit} ctas={-(-e.dims.npt//(4*mt))*nsplit}: {t0.elapsed_time(t1)/500*1e3:.1f} us/iter  {nb*500/(t0.elapsed_time(t1)*1e-3)/1e6:.1f} M problem-iters/s")
