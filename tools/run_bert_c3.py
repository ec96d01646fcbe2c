"""BERT4Rec at BASELINE configs[2] shape (L=200, d=256, H=4, |I|=100K, untied head): a few training steps, timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from replay_b200.engine_bert import Bert4RecEngine, BertConfig
from replay_b200.models.nn.sequential import uniform_masker
from replay_b200.synthetic import make_sequences
import sys as _s
B, L, d, I = (int(_s.argv[1]) if len(_s.argv) > 1 else 128), 200, 256, 100_000
cfg = BertConfig(n_items=I, d=d, n_heads=4, n_blocks=2, max_len=L, dropout=0.1)
eng = Bert4RecEngine(cfg, B, L, "cuda", seed=1)
ids, pm, _, _ = make_sequences(B, I, L, seed=3, pad_value=0)
tok = uniform_masker(pm, 0.15, torch.Generator().manual_seed(0))
ids, pm, tok = ids.cuda(), pm.cuda(), tok.cuda()
eng.set_batch(ids, pm, tok, ids)
losses = []
for i in range(8):
    losses.append(float(eng.train_step()[0]))
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 10
for i in range(n):
    eng.train_step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
from replay_b200.trainer import Trainer
tr = Trainer(eng)
for i in range(6):
    tr.step(ids, pm, tok, ids)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(50):
    loss = tr.step(ids, pm, tok, ids)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
print(f"bert4rec c3 under a CUDA graph: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} seq/s, loss {float(loss[0]):.3f}")
print("bert4rec c3: n_valid", int(eng.n_valid), "losses", [round(x, 3) for x in losses], f"eager {dt*1e3:.2f} ms/step -> {B/dt:.0f} seq/s", "launches/step", eng.lib.count // 18)
