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

x = e.gpart.view(torch.int64)[1024:1024 + 13].cpu().numpy()
xn = ["tile entry", "done/mu/slot loaded", "rhs formed", "Ginv gemm done", "KKT done", "PtP gemm done", "plane 0 done", "plane 1 done",
      "owner block entry", "operands staged", "tiles done", "threadfence done", "syncthreads done"]
print("x-update of CTA 0 / warp 0 (us after owner block entry):")
for i in (8, 9, 0, 1, 2, 3, 4, 5, 6, 10, 12):
    print(f"  {xn[i]:24s} {(x[i] - x[8]) / 1e3:7.2f}")

n = e.dims.nbal
raw = e.gpart.view(torch.int64)[2048:2048 + 2 * n].cpu().numpy().reshape(n, 2)
smid = raw[:, 0] & 4095
pc = np.stack([raw[:, 0] >> 12, raw[:, 1]], axis=1) / 1e3
own = pc[:, 0] > 0
print("CTAs per SM: ", np.bincount(np.bincount(smid[own])).tolist(), " smid of CTAs 0,1,2,148,149,296:", smid[[0, 1, 2, 148, 149, 296]].tolist())
print(f"owners: {own.sum()}  x-update done (us after entry): min {pc[own, 0].min():.2f} median {np.median(pc[own, 0]):.2f} max {pc[own, 0].max():.2f}")
print(f"first segment start: min {pc[:, 1].min():.2f} median {np.median(pc[:, 1]):.2f} max {pc[:, 1].max():.2f}")
print("owner done by CTA index (every 8th owner):", np.round(pc[own, 0][::8], 1).tolist())
print("segment start by CTA index (every 24th):", np.round(pc[::24, 1], 1).tolist())

xs = e.gpart.view(torch.int64)[4096:4096 + 16 * n].cpu().numpy().reshape(n, 16)[own]
rel_ = (xs - xs[:, 8:9]) / 1e3
print("per-owner x-update phases (us after owner block entry): min / median / max over the owners")
for i in (9, 1, 2, 3, 4, 5, 6, 10, 12):
    print(f"  {xn[i]:24s} {rel_[:, i].min():7.2f} {np.median(rel_[:, i]):7.2f} {rel_[:, i].max():7.2f}")

allx = e.gpart.view(torch.int64)[4096:4096 + 16 * n].cpu().numpy().reshape(n, 16)
cd, td, ent = allx[:, 13] / 1e3, allx[:, 14] / 1e3, allx[:, 15] / 1e3
print(f"all CTAs: chunks+epilogue done (us after own entry) min {cd.min():.2f} median {np.median(cd):.2f} max {cd.max():.2f};"
      f" tail done min {td.min():.2f} median {np.median(td):.2f} max {td.max():.2f}; entry spread {ent.max() - ent.min():.2f} us")
print("chunk phase length (segment start -> done): min %.2f median %.2f max %.2f" % tuple(np.percentile(cd - pc[:, 1], [0, 50, 100])))

T = (-(-e.dims.npt // (4 * e.dims.mt))) * (e.dims.nrt // 4)
nctc = e.dims.nrt // 4
gb = np.arange(n) * T // n
ge = (np.arange(n) + 1) * T // n
strad = (gb // nctc) != ((ge - 1) // nctc)
ln = cd - pc[:, 1]
print("chunk phase by kind: straddlers %d: median %.2f max %.2f | others: median %.2f max %.2f" % (strad.sum(), np.median(ln[strad]), ln[strad].max(), np.median(ln[~strad]), ln[~strad].max()))
print("by #chunks:", {int(k): round(float(np.median(ln[(ge - gb) == k])), 2) for k in np.unique(ge - gb)})
per_sm = {}
for i in range(n):
    per_sm.setdefault(int(smid[i]), []).append(ln[i])
sm_med = np.array([np.mean(v) for v in per_sm.values()])
sm_spread = np.array([max(v) - min(v) for v in per_sm.values()])
print("per-SM mean chunk phase: min %.2f median %.2f max %.2f; within-SM spread median %.2f max %.2f" % (sm_med.min(), np.median(sm_med), sm_med.max(), np.median(sm_spread), sm_spread.max()))
order = np.argsort(ln)[-8:]
print("slowest CTAs:", [(int(i), int(smid[i]), bool(strad[i]), int(ge[i] - gb[i]), round(float(pc[i, 1]), 1), round(float(ln[i]), 1)) for i in order])
