"""Per-phase timeline of the fused balanced step kernel (cfg3): build a traced copy of the library
(-DSPM_TRACE) and print the %globaltimer stamps of CTA 0 (an owner), CTA 1 and the last CTA for a few
consecutive launches inside a CUDA-graph replay:   python tools/pass_trace.py [nb]
    build:  nvcc ... -DSPM_TRACE -> /tmp/libadmm_trace.so   (done by this script)"""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = os.path.join(ROOT, "tools", "libadmm_trace.so")
csrc = os.path.join(ROOT, "admmsolver_b200", "csrc")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DSPM_TRACE", "-shared",
       "-o", lib] + [os.path.join(csrc, f) for f in ("primitives.cu", "spm.cu", "bp.cu", "peer.cu")] + ["-lcudart"]
if not os.path.exists(lib):
    subprocess.run(cmd, check=True)
os.environ["ADMM_B200_LIB"] = lib
import numpy as np
import torch
from admmsolver_b200 import batch, problems

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
basis = problems.ir_basis()
p = problems.spm_batch(min(nb, 4096), basis, Nw=2000, seed=1000)
g = torch.from_numpy(np.tile(p.g, (1, -(-nb // p.g.shape[1])))[:, :nb].copy()).cuda()
e = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb), g, lam=p.lam, mu=p.mu, batch_wide=True)
print(f"nb={nb} nsplit={e.dims.nsplit} nbal={e.dims.nbal} mt={e.dims.mt} step_mode={e._step_mode}")
e.solve(100, use_solo=False)
e.solve(100, use_solo=False)
torch.cuda.synchronize()
t = e.gpart.view(torch.int64)[:8 * 3 * 8].cpu().numpy().reshape(8, 3, 8)
last = int(e.lazy[3])
names = ["after pdl wait", "head done", "x-update/owner done", "segment start (x0 loaded)", "chunks done", "epilogue done", "tail done"]
order = [(last - k) & 7 for k in range(5, -1, -1)]
base = t[order[0], 0, 7]
for L in order:
    print(f"launch slot {L}:")
    for ci, cn in enumerate(("CTA 0", "CTA 1", "last CTA")):
        row = t[L, ci]
        print(f"  {cn:8s} entry {(row[7] - base) / 1e3:8.2f} us | " + "  ".join(f"{(row[i] - row[7]) / 1e3:6.2f}" for i in range(7)))
print("columns (us after kernel entry):", ", ".join(names))
