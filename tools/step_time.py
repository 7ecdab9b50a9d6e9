"""Time per iteration of the default launch configuration (python tools/step_time.py nb [nb ...]); ADMM_B200_LIB selects the build.
SYMMETRIC=1: sampling matrix with the exact parity of the IR basis (folded pass); SPLIT=1: additionally the x-update kernel +
pass kernel pair instead of the fused step (e._step_mode = 0)."""
import os
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from admmsolver_b200 import batch, problems
basis = problems.ir_basis()
p = problems.spm_batch(4096, basis, Nw=2000, seed=1000, symmetric=bool(os.environ.get("SYMMETRIC")))
for nb in [int(a) for a in sys.argv[1:]] or [262144]:
    g = torch.from_numpy(np.tile(p.g, (1, -(-nb // 4096)))[:, :nb].copy()).cuda()
    e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True)
    for mode in ([None, 0] if os.environ.get("SPLIT") else [None]):
        if mode is not None:
            e._step_mode = mode
        e.solve(30); torch.cuda.synchronize()
        best = 1e30
        for rep in range(3):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            e.solve(100)
            t1.record(); torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1) / 100 * 1e3)
        print(f"nb={nb} fold={e.fold} step_mode={e._step_mode} mt={e.dims.mt} nsplit={e.dims.nsplit} nbal={e.dims.nbal}: {best:.1f} us/iter "
              f"{nb / best:.2f} M problem-iters/s frac {nb * 324158.0 / (best * 1e-6) / 35.4e12:.4f}", flush=True)
    del e, g
