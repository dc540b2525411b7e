#!/usr/bin/env python
"""Where does the e2e gap go?  Times FullGraphTrainer.step with device vs pinned-host inputs (synchronised every
step, wall clock) and the bare H2D copy, on the bench atlas."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
from scdeepsort_b200.trainer import FullGraphTrainer

dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 760_000
g = synthetic_bipartite(cells, 20_000, 2000, device=dev)
feats = synthetic_features(g, 400)
labels = torch.randint(0, 16, (cells,)).to(dev)
g.densify(0.3)
tr = FullGraphTrainer(g, 16, dense_dim=400, hidden_dim=400, n_layers=2)
hf, hl = feats.cpu().pin_memory(), labels.cpu().pin_memory()
print("pinned:", hf.is_pinned(), hf[20000:].is_pinned())

def wall(fn, n=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n

buf = torch.empty_like(feats)
print("bare H2D copy ms:", wall(lambda: buf.copy_(hf, non_blocking=True)))
print("step(device), sync each step ms:", wall(lambda: tr.step(feats, labels)))
print("step(device), no per-step sync ms:", wall(lambda: tr.step(feats, labels, return_loss=False)))
print("step(host pinned) ms:", wall(lambda: tr.step(hf, hl)))
# forward only, host inputs: time to the end of the first aggregate
def fwd_host():
    with torch.no_grad():
        f, l, ready = tr._stage(hf, hl)
        torch.cuda.current_stream().wait_event(ready)
print("stage + wait ms:", wall(fwd_host))
