"""Sweep of the balanced decomposition (fused balanced step) over tiles per warp and number of pieces:
   [SYMMETRIC=1] python tools/bal_sweep.py [nb ...]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from admmsolver_b200 import batch, problems
basis = problems.ir_basis()
nbs = [int(a) for a in sys.argv[1:]] or [4096]
for nb in nbs:
    # SYMMETRIC=1: sampling matrix with the exact parity of the IR basis (folded pass)
    p = problems.spm_batch(min(nb, 4096), basis, Nw=2000, seed=1000, symmetric=bool(__import__("os").environ.get("SYMMETRIC")))
    g = np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb]
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True)
    import os
    if os.environ.get("SKEWS"):
        for sk in [float(v) for v in os.environ["SKEWS"].split(",")]:
            e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, bal_skew=sk)
            e.solve(100, use_solo=False); torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(5): e.solve(100, use_solo=False)
            t1.record(); torch.cuda.synchronize()
            us = t0.elapsed_time(t1) / 500 * 1e3
            print(f"nb={nb} skew={sk} mt={e.dims.mt} nsplit={e.dims.nsplit} nbal={e.dims.nbal}: {us:.1f} us/iter frac {nb * 324158.0 / (us * 1e-6) / 35.4e12:.3f}", flush=True)
        continue
    cfgs = [(None, None)] + [(mt, nbal) for mt in (1, 2) for nbal in ((148, 222, 296, 370, 444) if nb > 1500 else (32, 64, 128, 192, 256, 344, 444))]
    for mt, nbal in cfgs:
        try:
            e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, mt=mt, nbal=nbal)
        except Exception as ex:
            print(f"nb={nb} mt={mt} nbal={nbal}: {type(ex).__name__} {ex}")
            continue
        e.solve(100, use_solo=False); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): e.solve(100, use_solo=False)
        t1.record(); torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / 500 * 1e3
        print(f"nb={nb} mt={mt} nbal={nbal} -> mt={e.dims.mt} nsplit={e.dims.nsplit} nbal={e.dims.nbal} step_mode={e._step_mode}: {us:.1f} us/iter "
              f"{nb / us:.1f} M problem-iters/s  frac {nb * 324158.0 / (us * 1e-6) / 35.4e12:.3f}", flush=True)
