"""CE(sum) loss and Adam step kernels against torch (train.py:34-36,82-85)."""
import pytest
import torch

import scdeepsort_b200 as sd
from scds_helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("m,k", [(1, 2), (37, 4), (5000, 11), (100000, 16), (64, 100)])
def test_cross_entropy_sum_matches_torch(m, k):
    g = torch.Generator().manual_seed(m + k)
    logits = (torch.randn(m, k, generator=g) * 3).to(DEV).requires_grad_(True)
    labels = torch.randint(0, k, (m,), generator=g).to(DEV)
    loss = sd.optim.cross_entropy_sum(logits, labels)
    (loss * 0.5).backward()
    ref_in = logits.detach().double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, labels, reduction="sum")
    (ref * 0.5).backward()
    assert abs(float(loss) - float(ref)) < 2e-6 * abs(float(ref)) + 1e-6
    assert rel_err(logits.grad.cpu(), ref_in.grad.cpu()) < 1e-6


def test_adam_matches_torch_over_steps():
    torch.manual_seed(0)
    shapes = [(402, 1), (400, 400), (400,), (16, 400), (16,)]
    ours = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    theirs = [p.detach().clone().requires_grad_(True) for p in ours]
    o1 = sd.optim.Adam(ours, lr=1e-3, weight_decay=5e-4)
    o2 = torch.optim.Adam(theirs, lr=1e-3, weight_decay=5e-4)
    for step in range(6):
        for a, b in zip(ours, theirs):
            gr = torch.randn_like(a) * (10.0 ** (step - 3))
            a.grad, b.grad = gr.clone(), gr.clone()
        o1.step(); o2.step()
        for a, b in zip(ours, theirs):
            assert rel_err(a.detach().cpu(), b.detach().cpu()) < 1e-6
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert set(sd1["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"}
    assert sd1["param_groups"][0]["lr"] == sd2["param_groups"][0]["lr"]
    assert rel_err(sd1["state"][1]["exp_avg_sq"].cpu(), sd2["state"][1]["exp_avg_sq"].cpu()) < 2e-6


def test_cross_entropy_out_of_range_label_poisons_loss_without_oob_read():
    """the -1 of a gene row in ``all_labels`` must not index the logits: the loss becomes NaN instead."""
    logits = torch.randn(64, 5, device="cuda:0", requires_grad=True)
    labels = torch.randint(0, 5, (64,), device="cuda:0")
    assert bool(torch.isfinite(sd.optim.cross_entropy_sum(logits, labels)))
    for bad in (-1, 5, 10 ** 9):
        lab = labels.clone()
        lab[17] = bad
        assert bool(torch.isnan(sd.optim.cross_entropy_sum(logits, lab)))
