"""Micro-benchmark of the two head kernels at BASELINE config sizes (CUDA events, L2 flushed by input size)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from replay_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    ts.sort()
    return ts[len(ts) // 2]

res = {}
g = torch.Generator(device="cuda").manual_seed(0)
# ---- predict head C4: 500K items, d=128
for B in (512, 4096, 32768):
    I, d, K, S = 500_000, 128, 10, 200
    hq = (torch.randn(B, d, device="cuda", generator=g) * 0.5).bfloat16()
    table = (torch.randn(I, d, device="cuda", generator=g) * 0.5).bfloat16()
    seen = torch.randint(0, I, (B, S), device="cuda", generator=g)
    ss = ops.seen_prepare(seen, I)
    ms = timeit(lambda: ops.score_topk(hq, table, K, ss))
    flops = 2.0 * B * I * d
    res[f"score_topk_B{B}"] = dict(ms=ms, users_per_s=B / ms * 1e3, tflops=flops / ms / 1e9, table_gbs=I * d * 2 / ms / 1e6)
# ---- CE head C2: 50K items, d=128
for T in (28672, 51200):
    I, d = 50_000, 128
    hc = torch.randn(T, d, device="cuda", generator=g).bfloat16()
    table = (torch.randn(I, d, device="cuda", generator=g) * 0.3).bfloat16()
    labels = torch.randint(0, I, (T,), device="cuda", generator=g).int()
    nv = torch.tensor([T], dtype=torch.int32, device="cuda")
    st = ops.CEHeadState(T, I, d, "cuda")
    d_hc = torch.zeros(T, d, device="cuda", dtype=torch.bfloat16)
    d_tab = torch.zeros(I + 1, d, device="cuda")
    ms_f = timeit(lambda: ops.ce_head_fwd(st, hc, table, labels, nv, d_hc=d_hc, n_valid_hint=int(nv.item())))
    ms_b = timeit(lambda: ops.ce_head_bwd(st, hc, table, labels, nv, d_hc, d_tab))
    fl = 2.0 * T * I * d
    res[f"ce_T{T}"] = dict(fwd_ms=ms_f, bwd_ms=ms_b, fwd_tflops=fl / ms_f / 1e9, bwd_tflops_credited=2 * fl / ms_b / 1e9,
                           bwd_tflops_executed=4 * fl / ms_b / 1e9)
print(json.dumps(res, indent=1))
