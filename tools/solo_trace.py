"""Phase timeline of the cluster-resident solve (needs a library built with -DSOLO_TRACE):
ADMM_B200_LIB=tools/lib_trace.so python tools/solo_trace.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admmsolver_b200 import batch, problems  # noqa: E402

p = problems.spm_single(problems.ir_basis(), Nw=2000)
e = batch.SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True)
e.solve(50)
t = e.gpart.view(torch.int64).cpu().numpy()[:4 * 2 * 16].reshape(4, 2, 16)
names = ["top", "S1", "step2", "S2", "step3/rows", "S3", "Vpartial", "cluster", "remote", "sync", "exch-ret", "decide"]
for it in range(4):
    for w in range(2):
        st = t[it, w, :12]
        print("it %d %s: " % (10 + it, "L-warp0 " if w == 0 else "row-warp"),
              "  ".join("%s +%d" % (names[i], st[i] - st[i - 1]) for i in range(1, 12)), " | total to next top:",
              (t[it + 1, w, 0] - st[0]) if it < 3 else "-")
