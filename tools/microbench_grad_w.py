import sys, torch
sys.path.insert(0, "/root/repo")
from scdeepsort_b200 import dense
g = torch.randn(780000, 400, device="cuda"); x = torch.randn(780000, 400, device="cuda")
g_hi, g_lo, _ = dense.split_tf32(g); x_hi, x_lo, _ = dense.split_tf32(x)
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("grad_w_tc ms:", t(lambda: dense.grad_w_tc(g_hi, g_lo, x_hi, x_lo)))
print("torch g.t() @ x ms:", t(lambda: g.t() @ x))
d = dense.grad_w_tc(g_hi, g_lo, x_hi, x_lo); r = (g.double().t() @ x.double())
print("rel err:", float((d.double() - r).abs().max() / r.abs().max()))
