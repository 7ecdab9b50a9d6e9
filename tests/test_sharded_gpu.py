"""GPU: parity of the SHARDED engine -- ``SharedSpM(batch_wide=True, group=...)``, the path the 2^20-problem
sweep runs at N > 1 (BASELINE config 5; reference semantics: packed batch-global norms,
/root/reference/src/admmsolver/optimizer.py:232-299).

  * world size 1 (in-process group): ``admm_spm_reduce_post`` + ``admm_spm_decide_peer`` (mailbox path, CUDA-graph
    replay and eager) and the NCCL variant ``admm_spm_reduce`` + ``all_reduce`` + ``admm_spm_decide`` against the
    oracle's packed solve: x, mu history, iteration count, residual history;
  * world size 2 on ONE GPU: two processes, each owning a contiguous slab (``sharding.shard_range``), exchange
    their sums through each other's CUDA-IPC mailbox (the same kernels and the same peer-mapped stores that
    cross NVLink on the 8-GPU box) -- against the reference's golden packed solve of the whole batch.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden, rel

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def world1(build_lib):
    """A one-rank process group on the GPU (NCCL; the mailbox bootstrap and the NCCL variant both run over it)."""
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(29700 + os.getpid() % 1000)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist.group.WORLD
    if created:
        dist.destroy_process_group()


@pytest.mark.parametrize("collective,use_graph", [("peer", True), ("peer", False), ("peer-classic", True), ("nccl", False)])
def test_sharded_engine_world1_vs_oracle(world1, ir_basis, collective, use_graph):
    from admmsolver_b200 import batch, problems
    from oracle import flat
    nb, niter, interval = 37, 260, 50
    p = problems.spm_batch(nb, ir_basis, Nw=200, seed=21)
    classic = collective == "peer-classic"          # mailbox path with the three-kernel iteration (reduce_post + decide_peer)
    collective = collective.split("-")[0]
    e = batch.SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True, group=world1, collective=collective)
    e.use_lazy = not classic
    assert e.collective == collective and (e._peer is not None) == (collective == "peer")
    n = e.solve(niter, interval_update_mu=interval, use_graph=use_graph)
    st = flat.spm_solve(p.s, p.P, p.C, p.D, p.g, p.lam, niter, mu=p.mu, interval_update_mu=interval)
    assert n == st.niter_done == niter
    assert rel(e.x0(), st.x0) < TOL and rel(e.x1(), st.x1) < TOL and rel(e.x2(), st.x2) < TOL
    assert float(e.mu10[0]) == st.mu10 and float(e.mu20[0]) == st.mu20
    assert len(e.primal_residual) == niter
    assert rel(e.primal_residual, st.primal) < 1e-8 and rel(e.dual_residual, st.dual) < 1e-8
    # identical to the unsharded engine (the sums are the same ten doubles, only their route differs)
    u = batch.SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True)
    u.solve(niter, interval_update_mu=interval, use_solo=False)
    assert rel(e.x0(), u.x0()) < 1e-13 and float(u.mu20[0]) == float(e.mu20[0])


def test_sharded_engine_world1_early_stop_and_resume(world1, ir_basis):
    """Batch-wide stopping test through the mailbox path: same stopping iteration as the oracle; a second solve
    continues (the sequence numbers of the mailbox run on across solve() calls)."""
    from admmsolver_b200 import batch, problems
    from oracle import flat
    p = problems.spm_batch(9, ir_basis, Nw=120, seed=5)
    e = batch.SharedSpM(p.s, p.P, p.C, p.D, p.g, lam=p.lam, mu=p.mu, batch_wide=True, group=world1)
    n1 = e.solve(3000, rtol=1e-5)
    st = flat.spm_solve(p.s, p.P, p.C, p.D, p.g, p.lam, 3000, mu=p.mu, rtol=1e-5)
    assert st.niter_done < 3000 and len(e.primal_residual) == st.niter_done
    assert rel(e.x0(), st.x0) < TOL
    e.solve(40, rtol=1e-14)
    st = flat.spm_solve(p.s, p.P, p.C, p.D, p.g, p.lam, 40, mu=p.mu, rtol=1e-14, state=st)
    assert rel(e.x0(), st.x0) < TOL and float(e.mu20[0]) == st.mu20


# ------------------------------------------------------------------ two ranks, one GPU
def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # host transport of the IPC handles only
    from admmsolver_b200 import batch
    from admmsolver_b200.sharding import shard_range
    g = golden("after_solve")
    nb = 6
    b0, b1 = shard_range(nb, rank, world)
    e = batch.SharedSpM(g["spm6_s"], g["spm6_P"], g["spm6_C"], np.ones(b1 - b0), g["spm6_g"][:, b0:b1], lam=float(g["spm6_lam"]),
                        mu=float(g["spm6_mu"]), batch_wide=True, group=dist.group.WORLD, keep_x_old=True)
    assert e._peer is not None and e._peer.world == world
    n = e.solve(130)
    out[rank] = (b0, b1, n, e.x0(), e.x0_old(), float(e.mu10[0]), float(e.mu20[0]), list(e.primal_residual),
                 list(e.dual_residual))
    e._peer.close()
    dist.destroy_process_group()


def _compute_mode():
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=compute_mode", "--format=csv,noheader", "-i", "0"],
                           stdout=subprocess.PIPE, text=True, timeout=20)
        return r.stdout.strip()
    except Exception:
        return "unknown"


@pytest.mark.timeout(600)
def test_two_ranks_one_gpu_vs_reference_golden(build_lib):
    mode = _compute_mode()
    if "Exclusive" in mode or "Prohibited" in mode:
        pytest.skip("GPU compute mode %s: a second process cannot share the device" % mode)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29900 + os.getpid() % 1000
    mp.spawn(_rank_main, args=(world, port, out), nprocs=world, join=True)
    g = golden("after_solve")
    nb = 6
    x0_ref = g["spm6_x0"].reshape(-1, nb)
    xo_ref = g["spm6_x_old0"].reshape(-1, nb)
    for rank in range(world):
        b0, b1, n, x0, xo, mu10, mu20, primal, dual = out[rank]
        assert n == 130 == len(g["spm6_primal"])
        assert rel(x0, x0_ref[:, b0:b1]) < TOL and rel(xo, xo_ref[:, b0:b1]) < 1e-9
        assert rel(primal, g["spm6_primal"]) < 1e-8 and rel(dual, g["spm6_dual"]) < 1e-8     # batch-GLOBAL residuals
    # both ranks took the same mu decisions as the reference: one more update_mu() gives its mu_after_update
    assert out[0][5:7] == out[1][5:7]
