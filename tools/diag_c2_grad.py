"""Diagnostic (GPU box): which kernel limits the gradient accuracy of the L2H400 Adipose 64-seed batch."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import scdeepsort_b200 as sd
from scdeepsort_b200 import dense
from scdeepsort_b200.gnn import NodeUpdate
from scds_helpers import adipose_inputs, seeded_state, rel_err

z = np.load(ROOT / "tests/golden/adipose.npz")
tag = sys.argv[1] if len(sys.argv) > 1 else "L2H400"
x, _, feats = adipose_inputs(z)
g = sd.DeepSortGraph.from_expression(x, features=feats).to("cuda:0")
state = seeded_state(z, tag, 400, 4, int(z["num_genes"]))
n_layers = int(z[f"{tag}/n_layers"])
seeds = torch.from_numpy(z[f"{tag}/grad_seeds"]).long()
labels = torch.from_numpy(z["labels"].astype(np.int64)).to("cuda:0")
for tc, tcw in ((True, True), (True, False), (False, False)):
    NodeUpdate.use_tensor_cores = tc
    dense.use_tc_grad_w = tcw
    m = sd.GNN(400, int(z[f"{tag}/hidden"]), 4, n_layers, int(z["num_genes"]), activation=torch.relu).to("cuda:0")
    m.load_state_dict(state); m.train()
    nf = next(iter(sd.NeighborSampler(g, len(seeds), g.number_of_nodes(), n_layers, 'in', seed_nodes=seeds)))
    nf.copy_from_parent()
    loss = sd.optim.cross_entropy_sum(m(nf), labels[nf.layer_parent_nid(-1)])
    loss.backward()
    print(f"tensor-core linear {tc}, tensor-core grad_w {tcw}: loss {float(loss):.6f} (ref {float(z[tag + '/loss']):.6f})")
    for name, p in m.named_parameters():
        gr = p.grad.detach().cpu()
        if f"{tag}/grad/{name}" in z.files:
            print(f"    {name:28s} rel_err {rel_err(gr, z[f'{tag}/grad/{name}']):.3e}")
        else:
            ref = torch.from_numpy(z[f"{tag}/grad_sample/{name}"]).double()
            d = (gr.reshape(-1)[::53].double() - ref).abs()
            i = int(d.argmax())
            print(f"    {name:28s} sample err {float(d.max() / gr.double().abs().max()):.3e} (at sample {i}: got {float(gr.reshape(-1)[::53][i]):.6e} ref {float(ref[i]):.6e}; "
                  f"max|g| {float(gr.abs().max()):.3e}) norm err {abs(float(gr.double().norm()) - float(z[f'{tag}/grad_norm/{name}'])) / float(z[f'{tag}/grad_norm/{name}']):.3e}")
