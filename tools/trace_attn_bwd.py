"""Diagnostic: per-CTA phase timeline of rp::attn_bwd_kernel (build with RP_NVCC_EXTRA=-DRP_ATTN_TRACE, run on the GPU box):
   RP_NVCC_EXTRA=-DRP_ATTN_TRACE python -m replay_b200.build --force && python tools/trace_attn_bwd.py"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from replay_b200._lib import lib
from replay_b200.engine import EncoderConfig, SasRecEngine
from replay_b200.synthetic import make_sequences

B, L, d, H, I = 512, 200, 128, 2, 50000
cfg = EncoderConfig(n_items=I, d=d, n_heads=H, n_blocks=2, max_len=L, dropout=0.2, variant="new")
eng = SasRecEngine(cfg, B, L, torch.device("cuda"), seed=1)
ids, pm, lab, tm = make_sequences(B, I, L, seed=3)
eng.set_batch(ids.cuda(), pm.cuda(), lab.cuda(), tm.cuda())
for _ in range(3):
    eng.train_step()
torch.cuda.synchronize()
buf = np.zeros(1024 * 32, dtype=np.uint64)
fn = lib().rp_debug_attn_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, buf.size) == 0
t = buf.reshape(1024, 32).astype(np.int64)
names = {0: "entry", 1: "stats done", 2: "sync"}
for s in range(8):
    names[3 + 3 * s] = f"s{s} S ready"; names[4 + 3 * s] = f"s{s} done"; names[5 + 3 * s] = f"s{s} acc ready"
names.update({28: "dK/dV drained", 29: "dQ drained", 30: "exit sync"})
rel = t - t[:, :1]
for cta in (0, 1, 147, 148, 500, 1023):
    ev = [(k, rel[cta, k]) for k in range(32) if t[cta, k] != 0 and k in names]
    print(f"CTA {cta}: " + "  ".join(f"{names[k]}={v}" for k, v in ev))
tot = rel[:, 30]
print("cycles entry->exit: mean", tot.mean(), "p10", np.percentile(tot, 10), "p90", np.percentile(tot, 90))
for k in sorted(names):
    v = rel[:, k][t[:, k] != 0]
    if v.size: print(f"{names[k]:>16s}: mean {v.mean():9.0f}")
