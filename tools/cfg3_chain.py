"""Where one cfg3 iteration (4096 problems, unfused path) spends its time: graph-replayed chains of the x-update kernel
alone, the pass kernel alone, and both, in stream order with programmatic dependent launch (python tools/cfg3_chain.py [nb])."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems
from admmsolver_b200._lib import call, stream

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
basis = problems.ir_basis()
p = problems.spm_batch(min(nb, 4096), basis, Nw=2000, seed=1000)
g = torch.from_numpy(np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb].copy()).cuda()
e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True)
e.solve(120, use_solo=False)
torch.cuda.synchronize()
dref, bref = C.byref(e.dims), C.byref(e.bufs)
print(f"nb={nb} nsplit={e.dims.nsplit} nbal={e.dims.nbal} mt={e.dims.mt}")


def chain(kinds, n=64):
    def body():
        for _ in range(n):
            for k in kinds:
                if k == "x":
                    call("admm_spm_xupdate_lazy", dref, bref, None, 1, stream())
                elif k == "p":
                    call("admm_spm_pass_lazy", dref, bref, None, stream())
                elif k == "X":
                    call("admm_spm_xupdate", dref, bref, stream())
                elif k == "P":
                    call("admm_spm_pass", dref, bref, 0, stream())
                elif k == "r":
                    call("admm_spm_reduce_decide", dref, bref, 0, stream())
    body()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode="thread_local"):
        body()
    gr.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        gr.replay()
    t1.record()
    torch.cuda.synchronize()
    e.lazy.zero_()
    return t0.elapsed_time(t1) / (5 * n) * 1e3


for kinds in ("x", "p", "xp", "X", "P", "XP", "XPr"):
    print(f"chain {kinds:4s}: {chain(kinds):7.2f} us per round", flush=True)
