"""Where a single-problem basis-pursuit solve spends its time: python tools/bp_time.py [M N niter]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admmsolver_b200 import _lib, batch, problems  # noqa: E402
from admmsolver_b200._lib import call, ptr, stream  # noqa: E402

M, N, niter = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (200, 1000, 1000)
A, y, xa = problems.basis_pursuit_instance(M, N, 10, 0)
e = batch.BatchedBasisPursuit(A, y, 1.0, 0.1)
z = torch.zeros(1, N, dtype=torch.float64, device="cuda")


def timed(fn, reps=5):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0, e0.elapsed_time(e1) * 1e-3))
    return min(o[0] for o in out) * 1e3, min(o[1] for o in out) * 1e3


def full():
    e.set_state(x0=z, x1=z, h=z, mu=1.0)
    e.solve(niter)


print("solve(%d) from zero: wall %.3f ms, device %.3f ms; iterations %d, mu %g" % ((niter,) + timed(full) + (int(e.iters[0]), float(e.mu[0]))))
e._fill(1e-12, 100)
bref, st = C.byref(e.bufs), stream()


def factor():
    e.need_factor.fill_(1)
    call("admm_bp_factor", bref, ptr(e.info), st)


print("factor (M=%d): wall %.3f ms, device %.3f ms" % ((M,) + timed(factor)))


def iterate_only():
    e.iters.zero_()
    e.done.zero_()
    call("admm_bp_iterate", bref, 99, st)       # iterations 1..99: no mu update in between


factor()
print("99 iterations, one launch: wall %.3f ms, device %.3f ms" % timed(iterate_only))
