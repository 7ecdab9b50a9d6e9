#!/usr/bin/env python
"""bench.py -- FP64 ADMM problem-iterations/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload spm_sweep|spm_cfg3|bp_cfg4|bp_cfg1|spm_cfg2]
    python bench.py --impl reference ...        # the reference's own CPU implementation, bounded sample

One "step" is one ``solve(niter)`` over the whole resident batch (``--niter`` ADMM iterations,
default 100 = one mu-update interval).  Default workload: BASELINE config 5, the 2^20-problem
complex128 SpM sweep sharing one basis, batch-sharded over the N ranks (strong scaling) with the
batch-wide stopping criterion (the residual sums are all-reduced every iteration by the kernels
themselves over peer-mapped NVLink memory; NCCL is the bootstrap and the barrier).

Before anything is timed every rank passes a PARITY GATE: the engine that is about to be timed (same
size, same launch configuration, same sharding) solves a batch of replicas of 64 problems and is compared
with the unmodified reference's packed solve of those 64 problems (norms scale by the replica count, so mu
history and stopping test are those of the small batch); the line carries the result in ``parity``.

Rank 0 prints ONE JSON line.  Besides the headline workload it carries, under ``also``, the same
measurement for the other batched BASELINE workloads (``bp_cfg4``: 65536 independent basis-pursuit
problems, sharded without communication; ``spm_cfg3``: 4096 problems sharing one A, N = 1 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fp64_admm_problem_iters_per_sec"
UNIT = "problem-iters/s"
# SpM batches: sampling matrix with the exact parity of the IR basis (problems.spm_batch(symmetric=True)); --no-fold: as
# the quadrature delivers it (parity to 1e-9 only), which keeps the engine on the unfolded pass
SYMMETRIC_P = True


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spm_sweep",
                    choices=["spm_sweep", "spm_cfg3", "bp_cfg4", "bp_cfg1", "spm_cfg2"])
    ap.add_argument("--nb", type=int, default=None, help="total number of problems (default: per workload)")
    ap.add_argument("--niter", type=int, default=None, help="ADMM iterations per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--per-problem", action="store_true", help="SpM: per-problem mu/stopping (no collective)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="sharded batch-wide criterion: in-kernel peer-memory all-reduce (default) or NCCL between kernels")
    ap.add_argument("--no-also", action="store_true", help="only the headline workload (skip the `also` block)")
    ap.add_argument("--also", default=None, help="comma-separated subset of the `also` entries (default: all)")
    ap.add_argument("--no-parity-gate", action="store_true")
    ap.add_argument("--no-fold", action="store_true",
                    help="SpM: sampling matrix as the quadrature delivers it (parity of the IR basis only to 1e-9): no folded pass")
    return ap.parse_args()


WORKLOADS = {
    # name: (default total nb, default niter, description)
    "spm_sweep": (1 << 20, 100, "cfg5: 2^20 SpM problems sharing one IR basis (L=39, Nw=2000), complex128"),
    "spm_cfg3": (4096, 100, "cfg3: 4096 SpM problems sharing one A (L=39, Nw=2000), complex128"),
    "bp_cfg4": (65536, 100, "cfg4: 65536 independent basis-pursuit problems, own 128x512 A each"),
    "bp_cfg1": (1, 1000, "cfg1: basis pursuit 200x1000, 10-sparse, single problem"),
    "spm_cfg2": (1, 1000, "cfg2: SpM single problem L=39, Nw=2000"),
}


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU baselines (the reference's own implementation on the host cores; bounded samples)
# ----------------------------------------------------------------------------------------------
def _problems():
    """admmsolver_b200/problems.py loaded BY PATH: the input generators are plain NumPy, and importing the product
    package would map libadmm_b200.so into the process that times the reference."""
    mod = sys.modules.get("_admm_problems")
    if mod is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_admm_problems", os.path.join(ROOT, "admmsolver_b200", "problems.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_admm_problems"] = mod
        spec.loader.exec_module(mod)
    return mod


def _import_reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "admmsolver")):
        sys.path.insert(0, ref)
        import admmsolver  # noqa: F401
        return "reference"
    return "port"


def cpu_spm_sample(nb_s: int, niter: int, Nw: int = 2000, seed: int = 0):
    """Packed PartialDiagonalMatrix formulation of the reference (batch-wide), nb_s problems."""
    problems = _problems()
    kind = _import_reference()
    basis = problems.ir_basis()
    p = problems.spm_batch(nb_s, basis, Nw=Nw, seed=seed, symmetric=SYMMETRIC_P) if nb_s > 1 else problems.spm_single(basis, Nw=Nw)
    if kind == "reference":
        from admmsolver.matrix import DiagonalMatrix, PartialDiagonalMatrix, identity
        from admmsolver.objectivefunc import ConstrainedLeastSquares, L1Regularizer, NonNegativePenalty
        from admmsolver.optimizer import Model, SimpleOptimizer
        L = p.s.size
        if nb_s > 1:
            rest = (nb_s,)
            lstsq = ConstrainedLeastSquares(1.0, PartialDiagonalMatrix(-DiagonalMatrix(p.s), rest), p.g.ravel(),
                                            PartialDiagonalMatrix(p.C, rest), p.D.astype(float))
            conds = [(0, 1, identity(L * nb_s), identity(L * nb_s)),
                     (0, 2, PartialDiagonalMatrix(p.P, rest), identity(Nw * nb_s))]
        else:
            lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), p.g, p.C, p.D)
            conds = [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]
        opt = SimpleOptimizer(Model([lstsq, L1Regularizer(p.lam, L * nb_s), NonNegativePenalty(Nw * nb_s)], conds),
                              mu=p.mu)
        t0 = time.perf_counter()
        opt.solve(niter)
        dt = time.perf_counter() - t0
    else:
        from oracle import flat
        t0 = time.perf_counter()
        flat.spm_solve(p.s, p.P, p.C, p.D, p.g, p.lam, niter, mu=p.mu)
        dt = time.perf_counter() - t0
    return nb_s * niter / dt, kind, f"{nb_s} packed problems x {niter} iterations (L={p.s.size}, Nw={Nw}), {dt:.2f} s"


def cpu_bp_sample(nb_s: int, niter: int, M: int, N: int, K: int):
    problems = _problems()
    kind = _import_reference()
    A, y, _ = problems.basis_pursuit_batch(nb_s, M, N, K, 0)
    t0 = time.perf_counter()
    if kind == "reference":
        from admmsolver.matrix import identity
        from admmsolver.objectivefunc import L1Regularizer, LeastSquares
        from admmsolver.optimizer import Model, SimpleOptimizer
        for b in range(nb_s):
            opt = SimpleOptimizer(Model([LeastSquares(1.0, A[b], y[b]), L1Regularizer(0.1, N)],
                                        [(1, 0, identity(N), identity(N))]))
            opt.solve(niter)
    else:
        from oracle import flat
        for b in range(nb_s):
            flat.bp_solve(A[b], y[b], 1.0, 0.1, niter)
    dt = time.perf_counter() - t0
    return nb_s * niter / dt, kind, f"{nb_s} problems {M}x{N} x {niter} iterations, sequential instances, {dt:.2f} s"


def _mp_worker(args):
    """One single-threaded reference process of the multi-process CPU baseline."""
    kind, a, b, c = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except ImportError:
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        t0 = time.time()
        if kind == "spm":
            v, k, _ = cpu_spm_sample(a, b, seed=c)
            units = a * b
        else:
            v, k, _ = cpu_bp_sample(a, b, 128, 512, 10)
            units = a * b
        return units, t0, time.time(), k


def cpu_multiprocess(kind: str, per_proc: int, niter: int):
    """The same sample as one single-threaded reference process per host core, all running at once (at these
    sizes one BLAS thread per process beats one process with many threads).  Returns (value, kind, sample, nproc)."""
    import multiprocessing as mp
    nproc = len(os.sched_getaffinity(0))
    with mp.get_context("spawn").Pool(nproc) as pool:
        res = pool.map(_mp_worker, [(kind, per_proc, niter, i) for i in range(nproc)])
    units = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    what = "packed problems" if kind == "spm" else "problems 128x512, sequential instances"
    return (units / wall, res[0][3],
            f"{nproc} single-threaded processes x {per_proc} {what} x {niter} iterations, {wall:.2f} s", nproc)


def cpu_baseline_for(workload: str, scale: int = 1):
    """Bounded sample of the workload on all host threads (torchrun pins OMP_NUM_THREADS=1: undo it).
    Returns (value, kind, sample, threads actually used by the BLAS pool).  ``scale`` multiplies the iterations of
    the sample: 1 for the repeated steps of the reference arm (a few seconds each), 4 for the one-off ``cpu_baseline``
    leg of our arm (10-30 s of CPU work)."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=len(os.sched_getaffinity(0))):
            one = _cpu_baseline_for(workload, scale) + (host_threads(),)
    except ImportError:
        one = _cpu_baseline_for(workload, scale) + (host_threads(),)
    # batched workloads: also one single-threaded process per core; the faster of the two is the baseline
    if workload in ("spm_sweep", "spm_cfg3", "bp_cfg4") and len(os.sched_getaffinity(0)) > 1 \
            and not os.environ.get("ADMM_BENCH_NO_MP"):
        try:
            many = (cpu_multiprocess("spm", 64, 20 * scale) if workload.startswith("spm")
                    else cpu_multiprocess("bp", 4, 200 * scale))
            if many[0] > one[0]:
                return (many[0], many[1], many[2] + " (one multi-threaded process: %.0f)" % one[0], many[3])
            return (one[0], one[1], one[2] + " (one single-threaded process per core: %.0f)" % many[0], one[3])
        except Exception as exc:                       # a sandbox without process spawning: keep the single-process figure
            return (one[0], one[1], one[2] + " (multi-process variant unavailable: %s)" % type(exc).__name__, one[3])
    return one


def _cpu_baseline_for(workload: str, scale: int = 1):
    if workload in ("spm_sweep", "spm_cfg3"):
        return cpu_spm_sample(1024, 10 * scale)
    if workload == "spm_cfg2":
        return cpu_spm_sample(1, 400 * scale)
    if workload == "bp_cfg4":
        return cpu_bp_sample(16, 200 * scale, 128, 512, 10)
    return cpu_bp_sample(1, 600 * scale, 200, 1000, 10)


def host_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return len(os.sched_getaffinity(0))


# ----------------------------------------------------------------------------------------------
# workload description shared by both arms (the driver compares the two `config` objects)
# ----------------------------------------------------------------------------------------------
UNIT_BYTES = {"spm": 16.0 * 2000, "bp_cfg4": 8.0 * 128 * 512 + 8.0 * 128 * 128, "bp_cfg1": 8.0 * 200 * 1000 + 8.0 * 200 * 200}


def workload_config(workload: str, nb_total: int, world: int, niter: int, per_problem: bool, collective: str):
    nb_def, niter_def, desc = WORKLOADS[workload]
    is_spm = workload.startswith("spm")
    nb_local = nb_total // world if nb_total >= world else nb_total
    per_unit = UNIT_BYTES["spm"] if is_spm else UNIT_BYTES[workload]
    if is_spm and not per_problem:
        crit = "batch-wide" + (" (in-kernel peer-memory all-reduce over NVLink)" if world > 1 and collective == "peer"
                               else " (NCCL all-reduce)" if world > 1 else "")
    else:
        crit = "per-problem"
    return {"workload": workload, "description": desc, "problems_total": nb_local * world, "problems_per_gpu": nb_local,
            "iterations_per_step": niter, "criterion": crit,
            **({"sampling_matrix": "exact parity of the IR basis, P[Nw-1-r,l] = (-1)^l P[r,l]" if SYMMETRIC_P else
                "parity of the IR basis to 1e-9 only (--no-fold)"} if is_spm and nb_total > 1 else {}),
            "l2": "working set per iteration >> L2 (126 MB)" if nb_local * per_unit > 2.5e8 else
                  "working set fits L2; no flush (the solver iterates on resident state by design)"}


# ----------------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    nb_def, niter_def, desc = WORKLOADS[args.workload]
    vals, sample, kind, cores = [], "", "port", 1
    for i in range(args.warmup + args.steps):
        v, kind, sample, cores = cpu_baseline_for(args.workload)
        if i >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.nb or nb_def, max(1, world), args.niter or niter_def,
                                  args.per_problem, args.collective),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # evidence that none of this repo's native code is on the reference arm's path
        "native_so_mapped": "libadmm_b200" in open("/proc/self/maps").read(),
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, seconds: float = 1.5):
    """cuBLAS DGEMM 8192^3 via torch.matmul: burst (best of 5) and sustained (back-to-back)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    reps = max(3, int(seconds / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    sustained = 2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    torch.cuda.empty_cache()
    return burst, sustained


def _rel(a, b):
    return float(np.linalg.norm(np.ravel(a - b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def reference_packed_spm(p, niter):
    """The checker of the parity gate: the unmodified reference (baseline/_ref; the oracle port when it is absent) on
    the packed batch `p` -- x0 (L, nb), mu10, mu20, iterations run.  CPU, rank 0, a second or two."""
    kind = _import_reference()
    nb = p.g.shape[1] if p.g.ndim == 2 else 1
    L, Nw = p.s.size, p.P.shape[0]
    if kind == "reference":
        from admmsolver.matrix import DiagonalMatrix, PartialDiagonalMatrix, identity
        from admmsolver.objectivefunc import ConstrainedLeastSquares, L1Regularizer, NonNegativePenalty
        from admmsolver.optimizer import Model, SimpleOptimizer
        if nb > 1:
            rest = (nb,)
            lstsq = ConstrainedLeastSquares(1.0, PartialDiagonalMatrix(-DiagonalMatrix(p.s), rest), p.g.ravel(),
                                            PartialDiagonalMatrix(p.C, rest), np.ones(nb))
            conds = [(0, 1, identity(L * nb), identity(L * nb)), (0, 2, PartialDiagonalMatrix(p.P, rest), identity(Nw * nb))]
        else:
            lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), np.ravel(p.g), p.C, np.array([1.0]))
            conds = [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]
        opt = SimpleOptimizer(Model([lstsq, L1Regularizer(p.lam, L * nb), NonNegativePenalty(Nw * nb)], conds), mu=p.mu)
        opt.solve(niter)
        return (opt.x[0].reshape(L, nb), float(opt._mu[1, 0]), float(opt._mu[2, 0]), len(opt._primal_residual), kind)
    from oracle import flat
    st = flat.spm_solve(p.s, p.P, p.C, np.ones(nb), p.g.reshape(L, nb), p.lam, niter, mu=p.mu)
    return (np.asarray(st.x0).reshape(L, nb), float(st.mu10), float(st.mu20), int(st.niter_done), kind)


def reference_bp(A, y, niter):
    kind = _import_reference()
    N = A.shape[1]
    if kind == "reference":
        from admmsolver.matrix import identity
        from admmsolver.objectivefunc import L1Regularizer, LeastSquares
        from admmsolver.optimizer import Model, SimpleOptimizer
        opt = SimpleOptimizer(Model([LeastSquares(1.0, A, y), L1Regularizer(0.1, N)], [(1, 0, identity(N), identity(N))]))
        opt.solve(niter)
        return opt.x[0].real.copy(), float(opt._mu[1, 0]), len(opt._primal_residual), kind
    from oracle import flat
    st = flat.bp_solve(A, y, 1.0, 0.1, niter)
    return np.asarray(st.x0).real.copy(), float(st.mu), int(st.niter_done), kind


def run_workload(args, workload, ctx, steps, warmup, with_clocks):
    """Set up, gate, time and describe ONE workload.  Returns the dict of its part of the JSON line (rank 0; other
    ranks get None)."""
    import torch
    import torch.distributed as dist
    from admmsolver_b200 import _lib, batch, problems

    world, rank, local, group = ctx["world"], ctx["rank"], ctx["local"], ctx["group"]
    hbm_peak, hbm_src, fp64_burst, fp64_sus = ctx["hbm_peak"], ctx["hbm_src"], ctx["fp64_burst"], ctx["fp64_sus"]
    nb_def, niter_def, desc = WORKLOADS[workload]
    main_wl = workload == args.workload
    nb_total = (args.nb if main_wl and args.nb else nb_def)
    niter = (args.niter if main_wl and args.niter else niter_def)
    is_spm = workload.startswith("spm")
    nb_local = nb_total // world if nb_total >= world else nb_total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def bcast(obj):
        if world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    h2d = d2h = 0
    parity = None
    if is_spm:
        basis = problems.ir_basis()
        Nw = 2000
        L = basis.size
        # synthetic spectra: rank r owns the contiguous batch slab [r*nb_local, (r+1)*nb_local)
        gen_nb = min(nb_local, 4096)
        p = problems.spm_batch(gen_nb, basis, Nw=Nw, seed=1000 + rank, symmetric=SYMMETRIC_P) if nb_total > 1 else \
            problems.spm_single(basis, Nw=Nw)
        g_small = p.g.reshape(L, -1).astype(np.complex128)
        reps = -(-nb_local // g_small.shape[1])
        g_host = torch.from_numpy(np.tile(g_small, (1, reps))[:, :nb_local].copy()).pin_memory()
        out_host = torch.empty(L, nb_local, dtype=torch.complex128).pin_memory()
        g_dev = g_host.to("cuda", non_blocking=True)
        batch_wide = not args.per_problem
        eng = batch.SharedSpM(p.s, p.P, p.C, np.ones(nb_local), g_dev, lam=p.lam, mu=p.mu,
                              batch_wide=batch_wide, group=group if batch_wide else None, collective=args.collective)

        # a handful of problems: the cluster-resident single-launch solve (admm_spm_solo); every step
        # starts from the zero state so that it really runs `niter` iterations
        solo = (nb_local <= eng.SOLO_MAX_NB and group is None
                and _lib.lib.admm_spm_solo_supported(C.byref(eng.dims)) != 0)

        # ---- parity gate: THIS engine against the reference's packed solve (before any timing)
        if not args.no_parity_gate:
            gate_iters = 120 if nb_total > 1 else 200
            nd = 64 if nb_local % 64 == 0 else (nb_local if nb_local <= 64 else 0)
            if nd > 0:
                pg = problems.spm_batch(nd, basis, Nw=Nw, seed=4242, symmetric=SYMMETRIC_P) if nb_total > 1 else p
                gg = pg.g.reshape(L, -1).astype(np.complex128)
                R = nb_local // nd
                gg_dev = torch.from_numpy(gg).cuda().repeat(1, R).contiguous()       # replica r of problem d: column r*nd+d
                eng.reset(g=gg_dev, mu=pg.mu)
                ran = eng.solve(gate_iters)
                x0 = eng.x0_device()
                first = x0[:, :nd]
                replicas_equal = bool((x0.view(L, R, nd) == first[:, None, :]).all())
                if batch_wide or nd == 1:
                    ref = bcast(reference_packed_spm(pg, gate_iters) if rank == 0 else None)
                    x_ref, mu10_ref, mu20_ref, n_ref, checker = ref
                    err = max(_rel(x0[:, r * nd:(r + 1) * nd].cpu().numpy(), x_ref) for r in sorted({0, R // 2, R - 1}))
                    mu_eq = float(eng.mu10[0]) == mu10_ref and float(eng.mu20[0]) == mu20_ref
                    it_eq = int(ran) == int(n_ref)
                else:
                    # per-problem criterion: every column is its own reference instance (sampled problems)
                    import copy
                    err, mu_eq, it_eq = 0.0, True, True
                    for dcol in (0, nd // 2, nd - 1):
                        one = copy.copy(pg)
                        one.g = gg[:, dcol:dcol + 1]
                        x_ref, mu10_ref, mu20_ref, n_ref, checker = bcast(reference_packed_spm(one, gate_iters) if rank == 0 else None)
                        col = (R - 1) * nd + dcol
                        err = max(err, _rel(x0[:, col].cpu().numpy(), x_ref[:, 0]))
                        mu_eq = mu_eq and float(eng.mu10[col]) == mu10_ref and float(eng.mu20[col]) == mu20_ref
                        it_eq = it_eq and int(eng.iters[col]) == int(n_ref)
                # whole columns per CTA (the fused step kernel): a replica must come out bit-identical wherever it sits;
                # the balanced small-batch decomposition cuts the sampling points of a tile group at position-dependent
                # places, so there the replicas agree to rounding only (reported, not required)
                whole_columns = eng.dims.nsplit == 1 and eng.dims.nbal == 0
                bad = allmax(0.0 if (err <= 1e-10 and mu_eq and it_eq and (replicas_equal or not whole_columns)) else 1.0)
                parity = {"max_rel_x0": allmax(err), "tolerance": 1e-10, "mu_equal": allmax(0.0 if mu_eq else 1.0) == 0.0,
                          "iterations_equal": allmax(0.0 if it_eq else 1.0) == 0.0,
                          "replicas_bit_identical": allmax(0.0 if replicas_equal else 1.0) == 0.0,
                          "ranks_checked": world, "checker": checker,
                          "what": "%d replicas x %d problems per rank, %d iterations, vs the packed solve of the %d problems "
                                  "on the host (norms scale by the replica count)" % (R, nd, gate_iters, nd)}
                if bad:
                    raise SystemExit("bench.py: PARITY GATE FAILED, nothing timed: %s" % json.dumps(parity))
                del gg_dev, x0
                eng.reset(g=g_dev, mu=p.mu)

        def step():
            if solo:
                eng.reset(mu=p.mu)
            ran = eng.solve(niter)
            if ran != niter:        # an early batch-wide stop would silently inflate problem-iters/s
                raise SystemExit("bench.py: solve() ran %d of %d iterations inside the timed region" % (ran, niter))

        def e2e_step():
            g_dev.copy_(g_host, non_blocking=True)
            eng.reset(g=g_dev, mu=p.mu)
            ran = eng.solve(niter)
            if ran != niter:
                raise SystemExit("bench.py: solve() ran %d of %d iterations inside the e2e region" % (ran, niter))
            out_host.copy_(eng.x0_device(), non_blocking=True)

        # the end-to-end leg of the batched workloads goes through the product's host pipeline (batch.SpMHostStream):
        # every step uploads its data from pinned memory and downloads its result, the copies overlap the solves
        host_pipe = batch.SpMHostStream(eng) if not solo else None

        def e2e_run(nsteps):
            for k in range(nsteps):
                ran = host_pipe.submit(g_host, out_host, niter, mu=p.mu, next_g_host=g_host if k + 1 < nsteps else None)
                if ran != niter:
                    raise SystemExit("bench.py: solve() ran %d of %d iterations inside the e2e region" % (ran, niter))
            host_pipe.join()

        h2d = g_host.numel() * 16
        d2h = out_host.numel() * 16
        # dominant kernel = one ADMM iteration of the batch.  Large batches: the fused step kernel
        # (x-update + pass); small batches (row-split path): the pass kernel alone.
        #   pass: both skinny GEMMs on the REAL plane only, 4 L Nw flop, 8 B read + 8 B written per
        #         sampling point (the imaginary half of the state is advanced in L-space)
        #   x-update (fused only): per plane Ginv*rhs and PtP*x0 = 2 * 2 L^2 flop; per plane 6 L-vectors read
        #         and 4 written, plus z and a of the imaginary plane (2 read, 2 written)
        fused = eng._step_mode != 0          # the x-update runs inside the step kernel (whole-column or balanced form)
        npl = eng.dims.nplanes
        folded = bool(getattr(eng, "fold", False))      # pairs of sampling points share the MMAs (admm_spm_dims.fold)
        flops_per_unit = (2.0 if folded else 4.0) * L * Nw + (4.0 * L * L * npl if fused else 0.0)
        bytes_per_unit = 16.0 * Nw + (8.0 * L * (10 * npl + (4 if npl == 2 else 0)) if fused else 0.0)
        survey_flops_per_unit = 8.0 * L * Nw + 4.0 * L * L       # both planes through both skinny GEMMs
        kernel_name = "spm_pass_kernel<%d,%d,0,%d%s>" % (eng.dims.Lp // 8, eng.dims.mt, npl if fused else 0,
                                                         (",bal" if eng._step_mode == 2 else "") + (",fold" if folded else ""))
        if solo:
            flops_per_unit = 4.0 * L * Nw + 4.0 * L * L * npl
            bytes_per_unit = 0.0          # P, the state and the factor stay in shared memory / registers
            kernel_name = "spm_solo_kernel<8>"
        bound = "tensor"
    else:
        solo = False
        if workload == "bp_cfg4":
            M, N, K = 128, 512, 10
        else:
            M, N, K = 200, 1000, 10
        # generate on the host in slabs (identical bits for the sampled oracle problems), tile beyond 256
        gen_nb = min(nb_local, 256)
        A_s, y_s, _ = problems.basis_pursuit_batch(gen_nb, M, N, K, seed0=rank * gen_nb)
        reps = -(-nb_local // gen_nb)
        A_dev = torch.from_numpy(A_s).cuda().repeat(reps, 1, 1)[:nb_local].contiguous()
        y_host = torch.from_numpy(np.tile(y_s, (reps, 1))[:nb_local].copy()).pin_memory()
        out_host = torch.empty(nb_local, N, dtype=torch.float64).pin_memory()
        y_dev = y_host.cuda()
        eng = batch.BatchedBasisPursuit(A_dev, y_dev, 1.0, 0.1)
        zeros = torch.zeros(nb_local, N, dtype=torch.float64, device="cuda")

        def reset_state():
            eng.set_state(x0=zeros, x1=zeros, h=zeros, mu=1.0)

        def check_ran():
            lo = int(eng.iters.min().item())
            if lo != niter:
                raise SystemExit("bench.py: a problem ran %d of %d iterations inside the timed region" % (lo, niter))

        # ---- parity gate: sampled problems of THIS engine against independent reference instances
        if not args.no_parity_gate:
            reset_state()
            eng.solve(niter)
            x_all = eng._x0
            idx = sorted({0, min(1, gen_nb - 1), gen_nb - 1})
            err, mu_eq, it_eq, checker = 0.0, True, True, "reference"
            for b in idx:
                xr, mur, nr, checker = reference_bp(A_s[b], y_s[b], niter)
                last = b + (reps - 1) * gen_nb if b + (reps - 1) * gen_nb < nb_local else b     # its last replica in the slab
                for col in sorted({b, last}):
                    err = max(err, _rel(x_all[col].cpu().numpy(), xr))
                    mu_eq = mu_eq and float(eng.mu[col]) == mur
                    it_eq = it_eq and int(eng.iters[col]) == nr
            bad = allmax(0.0 if (err <= 1e-10 and mu_eq and it_eq) else 1.0)
            parity = {"max_rel_x0": allmax(err), "tolerance": 1e-10, "mu_equal": allmax(0.0 if mu_eq else 1.0) == 0.0,
                      "iterations_equal": allmax(0.0 if it_eq else 1.0) == 0.0, "ranks_checked": world, "checker": checker,
                      "what": "problems %s of every rank's slab (and their last replicas), %d iterations, vs independent "
                              "instances on the host" % (idx, niter)}
            if bad:
                raise SystemExit("bench.py: PARITY GATE FAILED, nothing timed: %s" % json.dumps(parity))

        def step():
            reset_state()
            eng.solve(niter)

        def e2e_step():
            # A stays resident (the operator); per step the data y travels in and x0 travels out
            y_dev.copy_(y_host, non_blocking=True)
            eng.set_data(y_dev)
            reset_state()
            eng.solve(niter)
            out_host.copy_(eng._x0, non_blocking=True)

        h2d = y_host.numel() * 8
        d2h = out_host.numel() * 8
        flops_per_unit = 4.0 * M * N + 2.0 * M * M
        survey_flops_per_unit = flops_per_unit
        if eng.At is not None:
            # single-sweep kernel: A streamed once per iteration (the column tile that yields A^T s also
            # feeds the next A r), K^-1 once; the N-vectors never leave shared memory
            bytes_per_unit = 8.0 * M * N + 8.0 * M * M
            kernel_name = "bp_fused_kernel<%d>" % (8 if M <= 128 else 16)
            if nb_local <= 16 and M <= 480 and N >= 128:
                # cluster-resident solve: A and K^-1 never leave shared memory; the same algorithmic bytes
                # divided by the time are an on-chip rate, not HBM traffic (latency-bound single problem)
                kernel_name = "bp_solo_kernel<16>"
        else:
            bytes_per_unit = 2.0 * 8 * M * N + 8.0 * M * M + 8.0 * 6 * N
            kernel_name = "bp_iterate_kernel"
        bound = "hbm"

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            fn()
        e1.record()
        barrier()
        return allmax(e0.elapsed_time(e1))

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
    launches0 = _lib.launch_count
    # SpM: iterations of >= ~1 ms are launched eagerly with CUDA events around the dominant kernel
    # inside the timed region; shorter ones are replayed from a CUDA graph (events cannot sit inside a
    # graph) and the kernel is timed in a separate eager pass over the same resident state.
    in_region_events = is_spm and nb_local >= 65536
    if in_region_events:
        eng.pass_events = []
    total_ms = timed(step, steps)
    launches = _lib.launch_count - launches0
    if not is_spm:
        check_ran()
    clocks = None
    if with_clocks:
        if total_ms < 1500.0:
            # short timed region: nvidia-smi (200 ms period) saw little of it; keep the identical load running,
            # untimed, until it has a few samples (same steps, same buffers)
            t_end = time.time() + 1.5
            while time.time() < t_end:
                step()
                torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None and total_ms < 1500.0:
            clocks["note"] = "sampled while the timed step kept running for 1.5 s after the %.0f ms timed region" % total_ms
    units = float(nb_local * world) * niter * steps
    value = units / (total_ms * 1e-3)

    # dominant-kernel duration (CUDA events on the launching stream)
    kernel_timing = None
    if is_spm and solo:
        k_ms = total_ms / steps
        units_per_launch = nb_local * niter
        kernel_timing = "step time (one cluster-resident launch runs all iterations; latency-bound, not a roofline case)"
    elif is_spm:
        if in_region_events:
            kernel_timing = "CUDA events around every launch inside the timed region"
        else:
            eng.pass_events = []
            step()
            barrier()
            kernel_timing = "separate eager pass after the timed region (timed region replays a CUDA graph)"
        durs = [a.elapsed_time(b) for a, b in eng.pass_events]
        eng.pass_events = None
        k_ms = float(np.mean(durs)) if durs else None
        units_per_launch = nb_local
    else:
        # the persistent kernel runs all iterations: time = step time / launches that do work
        k_ms = total_ms / steps
        units_per_launch = nb_local * niter
        kernel_timing = "step time (one persistent launch runs all iterations)"

    e2e = None
    if not args.no_e2e:
        if is_spm and host_pipe is not None:
            e2e_run(max(1, min(warmup, 2)))
            ms = timed(lambda: e2e_run(steps), 1)
        else:
            for _ in range(max(1, min(warmup, 2))):
                e2e_step()
            ms = timed(e2e_step, steps)
        e2e = {"value": units / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    if is_spm and eng._peer is not None:
        eng._peer.close()
    del eng
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of the
    # same kernel at the same per-launch size (profiles/traffic.json); null when no capture matches
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(kernel_name, {}).get(str(units_per_launch))
        if ent:
            traffic = float(ent["dram_bytes_read"]) + float(ent["dram_bytes_write"])
        else:
            ent = tj.get(kernel_name, {}).get("per_unit")
            if ent:     # capture of a smaller launch, traffic proportional to the units of the launch
                traffic = (float(ent["dram_bytes_read"]) + float(ent["dram_bytes_write"])) * units_per_launch
    except Exception:
        pass

    roof = None
    if k_ms:
        if bound == "tensor":
            ach = flops_per_unit * units_per_launch / (k_ms * 1e-3) / 1e12
            peak = fp64_sus
            roof = {"bound": "tensor", "kernel": kernel_name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": traffic,
                    "peak_source": "measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained (burst %.1f)" % fp64_burst,
                    "avg_launch_ms": k_ms, "kernel_timing": kernel_timing,
                    "algorithmic_flops_per_launch": flops_per_unit * units_per_launch,
                    "flops_note": "executed flops: %.0f kflop per problem-iteration (the imaginary plane is advanced in L-space%s, "
                                  "DESIGN.md 2.1); SURVEY 8(d) counts %.0f kflop for both planes through both skinny GEMMs -- "
                                  "the fraction is of EXECUTED flops" % (flops_per_unit / 1e3,
                                                                         "; pairs of sampling points share the MMAs" if is_spm and folded else "",
                                                                         survey_flops_per_unit / 1e3),
                    "iteration_frac": flops_per_unit * nb_local * niter * steps / (total_ms * 1e-3) / 1e12 / peak,
                    "algorithmic_bytes_per_launch": bytes_per_unit * units_per_launch,
                    "hbm_achieved_gbs": bytes_per_unit * units_per_launch / (k_ms * 1e-3) / 1e9,
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
                    "hbm_frac": bytes_per_unit * units_per_launch / (k_ms * 1e-3) / 1e9 / hbm_peak}
            if is_spm and folded and not solo:
                # a SPEED comparison with the unfolded kernel of earlier lines, not a utilisation: the same
                # problem-iterations per second priced at the unfolded algorithm's 4 L Nw + 8 L^2 flops
                unf = flops_per_unit + 2.0 * L * Nw
                roof["unfolded_equivalent"] = {"kflop_per_problem_iter": unf / 1e3,
                                               "kernel_frac_of_dgemm": unf * units_per_launch / (k_ms * 1e-3) / 1e12 / peak,
                                               "iteration_frac_of_dgemm": unf * nb_local * niter * steps / (total_ms * 1e-3) / 1e12 / peak,
                                               "note": "speed relative to the unfolded pass (which executes this many flops), NOT pipe utilisation"}
            if roof["hbm_frac"] > roof["frac"] and units_per_launch * bytes_per_unit > 2.5e8:      # (state streamed from HBM, not L2-resident)
                # (folded pass: half the tensor work per byte of state -- the kernel now sits closer to the HBM roof than
                # to the tensor roof; report the nearer one and keep the tensor figures beside it)
                roof.update({"bound": "hbm", "achieved": roof["hbm_achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                             "frac": roof["hbm_frac"], "peak_source": hbm_src,
                             "tensor_achieved_tflops": ach, "tensor_peak_tflops": peak, "tensor_frac": ach / peak,
                             "tensor_peak_source": "measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained (burst %.1f)" % fp64_burst})
        else:
            ach = bytes_per_unit * units_per_launch / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": kernel_name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": hbm_src, "avg_launch_ms": k_ms,
                    "kernel_timing": kernel_timing,
                    "algorithmic_bytes_per_launch": bytes_per_unit * units_per_launch,
                    "fp64_tflops": flops_per_unit * units_per_launch / (k_ms * 1e-3) / 1e12,
                    "fp64_peak_tflops": fp64_sus}

    cpu = None
    if not args.no_cpu_baseline:
        v, kind, sample, cores = cpu_baseline_for(workload, scale=4)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    return {
        "value": value, "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "scaling": "strong" if nb_total >= world else "weak",
        "config": workload_config(workload, nb_total, world, niter, args.per_problem, args.collective),
        "parity": parity, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }


def run_ours(args):
    global SYMMETRIC_P
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    fp64_burst = fp64_sus = None
    if rank == 0:
        fp64_burst, fp64_sus = measure_fp64_peak(torch)
    ctx = {"world": world, "rank": rank, "local": local, "group": group,
           "hbm_peak": float(peaks.get("hbm_gbs", 6650.0)),
           "hbm_src": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
           "fp64_burst": fp64_burst, "fp64_sus": fp64_sus}

    main_part = run_workload(args, args.workload, ctx, args.steps, args.warmup, with_clocks=True)
    also = {}
    if not args.no_also and args.workload == "spm_sweep" and args.nb is None:
        # the other batched BASELINE workloads, measured the same way in the same run (shorter: 3 warm-up, <= 3 steps)
        # (the single-GPU configurations of BASELINE.json -- cfg3 and the two single-problem cases cfg1, cfg2 -- at N = 1 only)
        pick = None if args.also is None else set(args.also.split(","))
        for wl in (["bp_cfg4"] + (["spm_cfg3", "bp_cfg1", "spm_cfg2"] if world == 1 else [])):
            if pick is not None and wl not in pick:
                continue
            part = run_workload(args, wl, ctx, min(args.steps, 3), 3, with_clocks=False)
            if part is not None:
                also[wl] = part
        if world == 1 and SYMMETRIC_P and (pick is None or "spm_sweep_general_P" in pick):
            # the headline sweep once more with the sampling matrix as the quadrature delivers it (parity of the IR basis
            # to 1e-9 only): the plain pass a general P gets.  Last, and guarded: it must never cost the main line.
            import copy
            a2 = copy.copy(args)
            a2.no_cpu_baseline = True
            SYMMETRIC_P = False
            try:
                part = run_workload(a2, "spm_sweep", ctx, min(args.steps, 3), 3, with_clocks=True)
                if part is not None:
                    also["spm_sweep_general_P"] = part
            except (Exception, SystemExit) as exc:      # (incl. the SystemExit of a failed gate)
                also["spm_sweep_general_P"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            finally:
                SYMMETRIC_P = True
    if rank == 0:
        line = {"metric": METRIC, "value": main_part["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_part["ms_per_step"], "higher_is_better": True,
                "scaling": main_part["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": main_part["config"], "parity": main_part["parity"], "roofline": main_part["roofline"],
                "cpu_baseline": main_part["cpu_baseline"], "e2e": main_part["e2e"], "gpu_launches": main_part["gpu_launches"],
                "clocks": main_part["clocks"], "fp64_peak_tflops": {"burst": fp64_burst, "sustained": fp64_sus}}
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global SYMMETRIC_P
    args = parse_args()
    SYMMETRIC_P = not args.no_fold
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
