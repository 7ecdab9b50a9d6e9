"""Wave quantisation of the fused step kernel: problem-iterations/s for the per-GPU batch sizes of the
1/2/4/8-GPU sweep with one or two problem tiles per warp (python tools/wave_sweep.py [nb ...])."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems

basis = problems.ir_basis()
p = problems.spm_batch(4096, basis, Nw=2000, seed=1000)
sizes = [int(a) for a in sys.argv[1:]] or [131072, 262144, 524288]
for nb in sizes:
    g = torch.from_numpy(np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb].copy()).cuda()
    for mt in (1, 2):
        e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, mt=mt, nsplit=1)
        e.solve(30)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            e.solve(60)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 180
        ctas = -(-e.dims.npt // (4 * mt))
        slots = 148 * (3 if mt == 2 else 4)
        print(f"nb={nb} mt={mt}: {ctas} CTAs = {ctas / slots:.2f} waves of {slots}: {ms * 1e3:.0f} us/iter, {nb / ms / 1e3:.1f} M problem-iters/s")
        del e
        torch.cuda.empty_cache()
