"""Time of the on-device IR basis (admm_svd_jacobi on the 1536 x 1312 weighted kernel) vs the host construction."""
import sys, time
sys.path.insert(0, ".")
import torch
from admmsolver_b200 import irbasis, problems
for eps in (1e-7, 1e-10):
    irbasis.ir_basis_device(eps=eps); torch.cuda.synchronize()
    t0 = time.perf_counter(); b = irbasis.ir_basis_device(eps=eps); torch.cuda.synchronize(); t1 = time.perf_counter()
    h = problems.ir_basis(eps=eps); t2 = time.perf_counter()
    print(f"eps={eps:g}: device L={b.size} sweeps={b.sweeps} {t1 - t0:.3f} s (incl. kernel samples on the host); host L={h.size} {t2 - t1:.3f} s")
