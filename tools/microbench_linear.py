"""Times the dense layer (forward, input gradient, weight gradient) at atlas shape on both implementations (GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from scdeepsort_b200 import dense

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 780_000
k = n = 400
x = torch.randn(rows, k, device="cuda")
w = (torch.randn(n, k, device="cuda") * 0.05).requires_grad_(True)
b = torch.zeros(n, device="cuda", requires_grad=True)
dy = torch.randn(rows, n, device="cuda")


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for impl in ("dense16", "tf32x3"):
    dense.use_dense16 = impl == "dense16"
    xr = x.clone().requires_grad_(True)
    t_f = timed(lambda: dense.linear_relu(xr, w, b, True))
    y = dense.linear_relu(xr, w, b, True)
    t_b = timed(lambda: torch.autograd.grad(y, (xr, w, b), dy, retain_graph=True))
    print(f"{impl}: forward {t_f:.3f} ms, backward (dx, dW, db) {t_b:.3f} ms   [{rows} x {k} -> {n}]", flush=True)
if dense.use_dense16 is False:
    dense.use_dense16 = True
fmt = dense._fmt()
a = dense._planes_a(x, fmt)
bw = dense._planes_b(w.detach(), fmt, k_is_row=False)
out = torch.empty(rows, n, device="cuda")
print(f"dense16 GEMM alone (planes ready): {timed(lambda: dense._gemm16(a, bw, fmt, 0, n, out, bias=b.detach(), relu=True)):.3f} ms", flush=True)
print(f"split BLOCKED(x): {timed(lambda: dense._planes_a(x, fmt)):.3f} ms;  KBLOCKS(x): {timed(lambda: dense._planes_b(x, fmt, k_is_row=True)):.3f} ms", flush=True)
am = dense._amax(x, fmt)
print(f"amax(x): {timed(lambda: dense._amax(x, fmt)):.3f} ms;  KBLOCKS(x) with amax given: {timed(lambda: dense._planes_b(x, fmt, k_is_row=True, amax=am)):.3f} ms "
      f"(ideal at 6.4 TB/s: 0.19 / 0.40 ms)", flush=True)
