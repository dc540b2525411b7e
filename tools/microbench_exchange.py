"""The exchange step of the cell-sharded pass, both forms, timed on the device (torchrun, one rank per GPU):

  nccl : wsage_sum_slabs -> ncclAllReduce -> dscale multiply -> self-loop addcmul      (4 launches + NCCL's)
  peer : wsage_peer_reduce                                                               (1 launch)

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/microbench_exchange.py
"""
import ctypes
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import scdeepsort_b200 as sd
from scdeepsort_b200 import _lib, peer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
G, D = 20_000, 400
slots = 20_224
iters = 50
lib = _lib.load()
out = {}
for n_slabs in (2, 4, 8):
    g = torch.Generator(device=dev).manual_seed(rank)
    slabs = torch.randn(n_slabs, slots, D, device=dev, generator=g)
    dscale = torch.rand(G, device=dev) + 0.5
    coef = torch.rand(G, device=dev)
    hg = torch.randn(G, D, device=dev)
    o1, o2 = torch.empty(G, D, device=dev), torch.empty(G, D, device=dev)
    raw1, raw2 = torch.empty(G, D, device=dev), torch.empty(G, D, device=dev)
    pg = peer.enable(G * D)
    assert pg is not None

    def nccl_form():
        # the product's NCCL path: slab sum (wsage_spmm's init path; here without the row map, as in peer_form), all-reduce, epilogue
        _lib.check(lib.wsage_sum_slabs(ctypes.c_void_p(slabs.data_ptr()), n_slabs, slots * D, G, D, ctypes.c_void_p(raw1.data_ptr()), D,
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "sum_slabs")
        dist.all_reduce(raw1)
        torch.mul(raw1, dscale[:, None], out=o1)
        o1.addcmul_(hg, coef[:, None])

    def peer_form():
        pg.reduce(slabs, G, dscale=dscale, selfcoef=coef, hself=hg, out=o2, raw=raw2)

    res = {}
    for name, fn in (("nccl", nccl_form), ("peer", peer_form)):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name + "_ms"] = float(t)
    pg.check()
    err = float((o1 - o2).abs().max() / o1.abs().max())
    res["rel_diff"] = err
    out[f"slabs{n_slabs}"] = res
if rank == 0:
    print(json.dumps({"world": world, "genes": G, "dim": D, "iters": iters, "note": "back-to-back calls, max over ranks; identity row map in both forms", **out}))
peer.disable()
dist.destroy_process_group()
