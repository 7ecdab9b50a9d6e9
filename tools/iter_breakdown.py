"""Per-iteration time of the SpM batch path at a per-GPU batch size, unsharded vs. the sharded (peer mailbox)
path with a one-rank group, CUDA-graph replay and eager:  python tools/iter_breakdown.py [nb ...]
Under `ncu --metrics gpu__time_duration.sum` with ADMM_BREAKDOWN_NCU=1 it only runs a few eager iterations."""
import os
import sys
import tempfile
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist
from admmsolver_b200 import batch, problems

dist.init_process_group("gloo", store=dist.FileStore(tempfile.mktemp(prefix="admm_bd_"), 1), rank=0, world_size=1)
basis = problems.ir_basis()
p = problems.spm_batch(4096, basis, Nw=2000, seed=1000)
sizes = [int(a) for a in sys.argv[1:]] or [131072, 4096]
ncu = bool(os.environ.get("ADMM_BREAKDOWN_NCU"))
for nb in sizes:
    g = torch.from_numpy(np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb].copy()).cuda()
    for name, kw in (("unsharded", {}), ("peer world-1", dict(group=dist.group.WORLD))):
        e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True, **kw)
        if ncu:
            e.solve(6, use_graph=False, use_solo=False)
            torch.cuda.synchronize()
            continue
        for graph in (True, False):
            e.solve(100, use_graph=graph, use_solo=False)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            n = 0
            for _ in range(3):
                n += e.solve(100, use_graph=graph, use_solo=False)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / n
            print(f"nb={nb} {name:13s} graph={graph}: {ms * 1e3:8.1f} us/iter  {nb / ms / 1e3:7.1f} M problem-iters/s  "
                  f"(nsplit={e.dims.nsplit} nbal={e.dims.nbal} mt={e.dims.mt})", flush=True)
        if e._peer is not None:
            e._peer.close()
        del e
        torch.cuda.empty_cache()
