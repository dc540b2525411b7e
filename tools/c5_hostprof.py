"""cProfile of the sampled (c5) training step's host side (GPU box): where the Python time goes when the step is launch-bound."""
import cProfile
import pstats
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import scdeepsort_b200 as sd
from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features

dev = torch.device("cuda:0")
cells, genes = (int(sys.argv[1]) if len(sys.argv) > 1 else 200_000), 20_000
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


def step():
    nf = next(it)
    nf.copy_from_parent()
    loss = sd.optim.cross_entropy_sum(model(nf), labels[nf.layer_parent_nid(-1)])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 30 * 1e3:.3f} ms per step (wall)")
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(40)
