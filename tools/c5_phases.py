"""Where the sampled (c5) step spends its time: sampler / copy_from_parent / forward / backward / optimiser (GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import scdeepsort_b200 as sd
from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features

dev = torch.device("cuda:0")
cells, genes = (int(sys.argv[1]) if len(sys.argv) > 1 else 760_000), 20_000
bg = synthetic_bipartite(cells, genes, 2000, device=dev)
feats = synthetic_features(bg, 400)
graph = sd.DeepSortGraph.from_bipartite(bg, feats)
del bg
labels = torch.cat([torch.full((genes,), -1, dtype=torch.int64), torch.randint(0, 16, (cells,))]).to(dev)
model = sd.GNN(400, 800, 16, 3, genes, activation=torch.relu).to(dev)
opt = sd.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
seeds = torch.arange(genes, genes + cells, device=dev)
sampler = sd.NeighborSampler(graph, 1024, num_hops=3, neighbor_type='in', shuffle=True, seed_nodes=seeds, fanouts=[25, 10, 5],
                             generator=torch.Generator(device=dev).manual_seed(1))
it = iter(sampler)
acc = dict(sample=0.0, copy=0.0, fwd=0.0, bwd=0.0, opt=0.0)
def tick():
    torch.cuda.synchronize(); return time.perf_counter()
for i in range(25):
    t0 = tick(); nf = next(it)
    t1 = tick(); nf.copy_from_parent()
    t2 = tick(); loss = sd.optim.cross_entropy_sum(model(nf), labels[nf.layer_parent_nid(-1)])
    t3 = tick(); opt.zero_grad(set_to_none=True); loss.backward()
    t4 = tick(); opt.step()
    t5 = tick()
    if i >= 5:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[k] += v * 1e3 / 20
print({k: round(v, 3) for k, v in acc.items()}, "ms per step (synchronised phases); sum", round(sum(acc.values()), 3))
