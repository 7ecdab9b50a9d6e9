"""Basis pursuit 128x512: cluster-resident solve (16-CTA cluster per problem, several waves) vs the fused streaming
kernel as a function of the batch size: ADMM_BP_SOLO_MAX=1000 python tools/bp_solo_nb_sweep.py"""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from admmsolver_b200 import batch, problems

A, y, _ = problems.basis_pursuit_batch(64, 128, 512, 10, seed0=0)
for nb in (4, 8, 16, 24, 32, 48, 64):
    for solo in (True, False):
        if solo:
            os.environ.pop("ADMM_BP_NO_SOLO", None)
        else:
            os.environ["ADMM_BP_NO_SOLO"] = "1"
        e = batch.BatchedBasisPursuit(A[:nb], y[:nb], 1.0, 0.1)
        z = torch.zeros(nb, 512, dtype=torch.float64, device="cuda")
        e.solve(200)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            e.set_state(x0=z, x1=z, h=z, mu=1.0)
            e.solve(400)
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / 1200 * 1e3
        print(f"nb={nb} solo={solo}: {us:.1f} us per iteration of the batch, {nb / us:.2f} M problem-iters/s")
