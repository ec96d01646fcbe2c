import os, sys, torch
sys.path.insert(0, "/root/repo")
from replay_b200 import ops
g = torch.Generator(device="cuda").manual_seed(5)
M, N, K = 819200, 256, 128
A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.2).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
def t(n=20):
    ops.gemm(A, W, C, M, N, K, bias=b); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): ops.gemm(A, W, C, M, N, K, bias=b)
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / n
os.environ.pop("RP_GEMM_WS_NO_WIDE", None); w = t(); cw = C[:1000].float().clone()
os.environ["RP_GEMM_WS_NO_WIDE"] = "1"; nw = t()
print("wide %.1f us  two-tile %.1f us  max diff %.4f  (HBM floor %.0f us)" % (w * 1e3, nw * 1e3, float((C[:1000].float() - cw).abs().max()), (M*K*2 + M*N*2) / 6.5e12 * 1e6))
