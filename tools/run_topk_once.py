import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from replay_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
I, d, K, S = 500_000, 128, 10, 200
g = torch.Generator(device="cuda").manual_seed(0)
hq = (torch.randn(B, d, device="cuda", generator=g) * 0.5).bfloat16()
table = (torch.randn(I, d, device="cuda", generator=g) * 0.5).bfloat16()
seen = torch.randint(0, I, (B, S), device="cuda", generator=g)
ss = ops.seen_prepare(seen, I)
for _ in range(2):
    ops.score_topk(hq, table, K, ss)
torch.cuda.synchronize()
