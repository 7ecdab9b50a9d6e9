"""CPU, world_size 2 over gloo: host-side logic of the sharded batch-wide criterion.

  * ``sharding.shard_range`` (contiguous slabs) and ``peer.exchange_handles`` / ``peer.check_one_box`` -- the
    PRODUCT's bootstrap that carries the CUDA-IPC mailbox handles between the ranks (any ``torch.distributed``
    backend; gloo here, NCCL on the GPU box) -- driven by two real processes;
  * the SEMANTICS the device path has to reproduce, stated with the oracle: each rank owns a slab and sums the ten
    squared-norm partials over the ranks every iteration; the result must equal the single-process packed solve
    (identical mu history, iteration count and x on every slab).  This second test exercises ``oracle/`` only; the
    product kernels (``admm_spm_reduce_post`` / ``admm_spm_decide_peer``) are tested against the same golden on the
    GPU by ``tests/test_sharded_gpu.py`` (world size 1 in-process and two ranks on one GPU).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from admmsolver_b200.sharding import shard_range
    from oracle import flat
    g = golden("spm_packed")
    nb = g["g"].shape[1]
    b0, b1 = shard_range(nb, rank, world)

    def allreduce(v):
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t)
        return t.numpy()

    st = flat.spm_solve(g["s"], g["P"], g["C"], np.ones(b1 - b0), g["g"][:, b0:b1], float(g["lam"]), 400,
                        mu=float(g["mu"]), allreduce=allreduce)
    out[rank] = (b0, b1, st.x0, st.mu10, st.mu20, st.niter_done, st.primal[-1])
    dist.destroy_process_group()


def _handle_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from admmsolver_b200 import peer
    import socket
    fake = bytes([rank]) * 64                      # stands for the 64-byte cudaIpcMemHandle_t of this rank's mailbox
    recs = peer.exchange_handles(dist.group.WORLD, (socket.gethostname(), os.getpid(), fake))
    peer.check_one_box(recs)
    out[rank] = [(h, pid, bytes(b)) for (h, pid, b) in recs]
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_handle_exchange_over_gloo(build_lib):
    """The mailbox bootstrap of ``peer.PeerMailbox``: every rank ends up with every rank's handle, in rank order."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29300 + os.getpid() % 1000
    mp.spawn(_handle_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] == out[1] and len(out[0]) == world
    for r, (host, pid, h) in enumerate(out[0]):
        assert h == bytes([r]) * 64
    assert len({pid for _, pid, _ in out[0]}) == world


def test_peer_one_box_checks(build_lib):
    from admmsolver_b200 import peer
    peer.check_one_box([("a", 1, b""), ("a", 2, b"")])
    with pytest.raises(NotImplementedError):
        peer.check_one_box([("a", 1, b""), ("b", 2, b"")])          # two hosts: no CUDA IPC
    with pytest.raises(NotImplementedError):
        peer.check_one_box([("a", 1, b""), ("a", 1, b"")])          # one process cannot map its own handle
    with pytest.raises(NotImplementedError):
        peer.check_one_box([("a", i, b"") for i in range(17)])


def test_shard_range_partition():
    from admmsolver_b200.sharding import shard_range
    for nb in (0, 1, 7, 8, 1 << 20, 1000003):
        for world in (1, 2, 4, 8):
            r = [shard_range(nb, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nb
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_oracle_sharded_semantics_equal_packed(build_lib):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    g = golden("spm_packed")
    nb = g["g"].shape[1]
    x0_ref = g["x0"].reshape(-1, nb)
    for rank in range(world):
        b0, b1, x0, mu10, mu20, nit, primal = out[rank]
        err = np.linalg.norm(x0 - x0_ref[:, b0:b1]) / np.linalg.norm(x0_ref[:, b0:b1])
        assert err < 1e-11
        assert (mu10, mu20) == (float(g["mu10"]), float(g["mu20"]))
        assert nit == len(g["primal"])
        assert abs(primal - g["primal"][-1]) / g["primal"][-1] < 1e-9
