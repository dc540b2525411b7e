"""wsage_peer_reduce (include/wsage.h): the exchange step of the cell-sharded pass over peer memory.

On one GPU the ranks are emulated inside one process (``PeerGroup.local_ranks``: every "rank" has its own peer allocation
and stream, the kernels spin on each other's flags exactly as across GPUs); with two or more GPUs the real thing runs:
two processes, cudaIpc-mapped buffers, a sharded training step against the NCCL all-reduce form of the same step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scdeepsort_b200 as sd
from scdeepsort_b200 import parallel, peer
from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
from scdeepsort_b200.trainer import FullGraphTrainer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _expected(slabs, slot, rows, dscale, selfcoef, hself):
    """Same association as the kernel: slabs in order, then ranks in order (fp32 adds are then bit-reproducible)."""
    parts = []
    for s in slabs:
        acc = s[0][slot.long()][:rows]
        for k in range(1, s.shape[0]):
            acc = acc + s[k][slot.long()][:rows]
        parts.append(acc)
    raw = parts[0]
    for p in parts[1:]:
        raw = raw + p
    out = raw.double() * dscale.double()[:, None] + selfcoef.double()[:, None] * hself.double()
    return raw, out


def _run_local(groups, slabs, slot, rows, dscale, selfcoef, hself):
    world = len(groups)
    dim = slabs[0].shape[2]
    outs = [torch.full((rows, dim), float("nan"), device=DEV) for _ in range(world)]
    raws = [torch.full((rows, dim), float("nan"), device=DEV) for _ in range(world)]
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for r in range(world):
        groups[r].reduce(slabs[r], rows, slot_of_row=slot, dscale=dscale, selfcoef=selfcoef, hself=hself[r], out=outs[r], raw=raws[r],
                         stream=streams[r])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    return outs, raws


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_local_ranks_match_the_ordered_sum_and_agree_bitwise(world):
    g = torch.Generator(device=DEV).manual_seed(world)
    max_rows, dim = 3000, 96
    groups = peer.PeerGroup.local_ranks(world, max_rows * dim, timeout_s=20.0)
    try:
        for call, rows in enumerate([3000, 1237, 5, 2999, 64]):              # odd and even calls: both buffer parities, reused
            slab_rows = 3072
            slot = torch.randperm(slab_rows, device=DEV, generator=g)[:rows].to(torch.int32)
            slabs = [torch.randn(1 + (r + call) % 4, slab_rows, dim, device=DEV, generator=g) for r in range(world)]
            dscale = torch.rand(rows, device=DEV, generator=g) + 0.5
            selfcoef = torch.randn(rows, device=DEV, generator=g)
            hself = torch.randn(rows + 3, dim + 4, device=DEV, generator=g)[:rows, :dim]        # a strided view
            outs, raws = _run_local(groups, slabs, slot, rows, dscale, selfcoef, [hself] * world)
            raw, out = _expected(slabs, slot, rows, dscale, selfcoef, hself)
            for r in range(world):
                assert torch.equal(raws[r], raw), (call, r)
                assert torch.equal(outs[r], outs[0]), (call, r)              # replicated gene state stays replicated
                assert float((outs[r].double() - out).abs().max()) <= 1e-6 * float(out.abs().max())
        for pg in groups:
            pg.check()
            assert pg.calls == 5
    finally:
        for pg in groups:
            pg.close()


def test_identity_row_map_and_outputs_are_optional():
    groups = peer.PeerGroup.local_ranks(2, 400 * 64)
    try:
        slabs = [torch.randn(2, 400, 64, device=DEV) for _ in range(2)]
        raws = [torch.empty(400, 64, device=DEV) for _ in range(2)]
        streams = [torch.cuda.Stream(device=DEV) for _ in range(2)]
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for r in range(2):
            groups[r].reduce(slabs[r], 400, raw=raws[r], stream=streams[r])
        torch.cuda.synchronize()
        want = (slabs[0][0] + slabs[0][1]) + (slabs[1][0] + slabs[1][1])
        assert torch.equal(raws[0], want) and torch.equal(raws[1], want)
        with pytest.raises(RuntimeError, match="neither out nor raw"):
            groups[0].reduce(slabs[0], 400)
        with pytest.raises(RuntimeError, match="max_elems"):
            groups[0].reduce(torch.randn(1, 401, 64, device=DEV), 401, raw=torch.empty(401, 64, device=DEV))
    finally:
        for pg in groups:
            pg.close()


def test_a_missing_rank_times_out_instead_of_hanging():
    groups = peer.PeerGroup.local_ranks(2, 128 * 32, timeout_s=0.3)
    try:
        slabs = torch.randn(1, 128, 32, device=DEV)
        out = torch.zeros(128, 32, device=DEV)
        groups[0].reduce(slabs, 128, raw=out)            # rank 1 never calls
        with pytest.raises(RuntimeError, match="timed out"):
            groups[0].check()
        groups[1].check()                                # rank 1 saw nothing
    finally:
        for pg in groups:
            pg.close()
    # the device is still healthy
    assert float(torch.ones(4, device=DEV).sum()) == 4.0


# ---- two real processes / GPUs -------------------------------------------------------------------------------------
C, G, DEG, D0, H, K = 6000, 700, 120, 64, 64, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _steps(rank, world, use_peer, n_steps=3):
    lo, hi = parallel.cell_ranges(C, world)[rank]
    dev = torch.device("cuda", rank)
    bg = synthetic_bipartite(C, G, DEG, device=dev, cell_range=(lo, hi)).densify(0.0)
    parallel.globalize_gene_normalisers(bg)
    feats = synthetic_features(bg, D0)
    labels = torch.randint(0, K, (hi - lo,), generator=torch.Generator().manual_seed(5 + rank)).to(dev)
    os.environ["WSAGE_PEER"] = "1" if use_peer else "0"
    tr = FullGraphTrainer(bg, K, dense_dim=D0, hidden_dim=H, n_layers=2, seed=3, sharded=True)
    assert (tr.peer_group is not None) == use_peer
    parallel.broadcast_params(tr.model)
    losses = [tr.step(feats, labels) for _ in range(n_steps)]
    if tr.peer_group is not None:
        tr.peer_group.check()
        assert tr.peer_group.calls == n_steps        # one gene-producing layer per step
    params = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()]).cpu().numpy()
    return losses, params


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        a = _steps(rank, world, use_peer=True)
        peer.disable()
        b = _steps(rank, world, use_peer=False)
        q.put((rank, a, b))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_two_processes_train_alike_with_the_peer_kernel_and_with_nccl():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (_, (l0, p0), (ln0, pn0)), (_, (l1, p1), (ln1, pn1)) = res
    assert np.array_equal(p0, p1)                    # replicas stay bit-identical under the peer kernel
    assert np.allclose(l0, ln0, rtol=1e-5) and np.allclose(l1, ln1, rtol=1e-5)
    assert float(np.abs(p0 - pn0).max()) < 1e-5 * float(np.abs(pn0).max())
