"""Phase timeline of the cluster-resident basis-pursuit solve (library built with -DSOLO_TRACE):
ADMM_B200_LIB=tools/lib_trace.so python tools/bp_solo_trace.py [M N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admmsolver_b200 import batch, problems  # noqa: E402

M, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 1000)
A, y, xa = problems.basis_pursuit_instance(M, N, 10, 0)
e = batch.BatchedBasisPursuit(A, y, 1.0, 0.1, keep_history=True)
e.solve(5000, interval_update_mu=10000)          # history: 5000 x 2 doubles; stamps sit behind the first 4096 int64
t = e.history.view(torch.int64).cpu().numpy().ravel()[4096:4096 + 64].reshape(4, 16)
names = ["top", "AT.s", "sync", "x-update", "sync", "A.r", "sync", "rs-push", "rs-wait", "sum+ag(t)-push", "sync", "decide", "ag(t)-wait+K.t+push",
         "allgather-wait"]
for it in range(4):
    st = t[it]
    print("it %d: " % (3000 + it), "  ".join("%s +%d" % (names[i], st[i] - st[i - 1]) for i in range(1, 14)),
          "| iteration:", (t[it + 1, 0] - st[0]) if it < 3 else "-")
