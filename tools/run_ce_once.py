import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from replay_b200 import ops
T, nv_, I, d = 102400, 55574, 50000, 128  # bench default: 512 sequences per step
g = torch.Generator(device="cuda").manual_seed(0)
hc = torch.randn(T, d, device="cuda", generator=g).bfloat16(); hc[nv_:] = 0
table = (torch.randn(I, d, device="cuda", generator=g) * 0.3).bfloat16()
labels = torch.randint(0, I, (T,), device="cuda", generator=g).int()
nv = torch.tensor([nv_], dtype=torch.int32, device="cuda")
st = ops.CEHeadState(T, I, d, "cuda")
d_hc = torch.zeros(T, d, device="cuda", dtype=torch.bfloat16); d_tab = torch.zeros(I + 1, d, device="cuda")
for _ in range(2):
    ops.ce_head_fwd(st, hc, table, labels, nv, d_hc=d_hc, n_valid_hint=int(nv.item()))
    ops.ce_head_bwd(st, hc, table, labels, nv, d_hc, d_tab)
torch.cuda.synchronize()
