"""Diagnostic: timeline of CTA 0 of the CE-head kernels (fused forward + dH pass, dE pass), first 256 column tiles.
Build a traced variant (tools: nvcc ... -DRP_CE_TRACE -o replay_b200/build/variants/ce_trace.so) and run on the GPU box:
   python tools/trace_ce.py replay_b200/build/variants/ce_trace.so"""
import ctypes, sys
import numpy as np, torch
P, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
T, nv_, I, d = 102400, 55574, 50000, 128
g = torch.Generator(device="cuda").manual_seed(0)
hc = torch.randn(T, d, device="cuda", generator=g).bfloat16(); hc[nv_:] = 0
table = (torch.randn(I, d, device="cuda", generator=g) * 0.3).bfloat16()
labels = torch.randint(0, I, (T,), device="cuda", generator=g).int()
nv = torch.tensor([nv_], dtype=torch.int32, device="cuda")
loss = torch.zeros(2, device="cuda"); lse = torch.zeros(T, device="cuda"); cvec = torch.full((T,), float("-inf"), device="cuda")
d_hc = torch.zeros(T, d, device="cuda", dtype=torch.bfloat16); d_tab = torch.zeros(I + 1, d, device="cuda")
st = torch.cuda.current_stream().cuda_stream
L = ctypes.CDLL(sys.argv[1])
L.rp_ce_head_workspace.restype = sz; L.rp_ce_head_workspace.argtypes = [ci, ci, ci]
L.rp_ce_head_fwd.argtypes = [P, P, P, P, P, ci, ci, ci, P, P, P, P, ci, P, sz, P]
L.rp_ce_head_bwd.argtypes = [P, P, P, P, P, ci, ci, ci, P, P, P, P, P, ci, ci, P, sz, P]
L.rp_debug_ce_trace.argtypes = [P, ci]
wsb = L.rp_ce_head_workspace(T, I, d); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
for _ in range(3):
    assert L.rp_ce_head_fwd(hc.data_ptr(), table.data_ptr(), None, labels.data_ptr(), nv.data_ptr(), T, I, d, loss.data_ptr(), lse.data_ptr(), cvec.data_ptr(), d_hc.data_ptr(), nv_, ws.data_ptr(), wsb, st) == 0
    assert L.rp_ce_head_bwd(hc.data_ptr(), table.data_ptr(), None, labels.data_ptr(), nv.data_ptr(), T, I, d, loss.data_ptr(), cvec.data_ptr(), d_hc.data_ptr(), d_tab.data_ptr(), None, 1, 0, ws.data_ptr(), wsb, st) == 0
torch.cuda.synchronize()
buf = np.zeros(2 * 16 * 256, dtype=np.uint64)
assert L.rp_debug_ce_trace(buf.ctypes.data, buf.size) == 0
t = buf.reshape(2, 16, 256).astype(np.int64)
names = ["epi arrives", "S seen", "exps done", "G handed", "mma waits G", "mma has G", "mma queued"]
for m, title in enumerate(("fused fwd+dH", "dE")):
    tt = t[m]; base = tt[0, 0]
    print(f"== {title}: events relative to the epilogue's arrival at tile 0 (cycles)")
    for j in list(range(0, 8)) + list(range(100, 108)):
        print(f"tile {j:3d}: " + "  ".join(f"{names[k]}={tt[k, j] - base}" for k in range(7)))
    dj = np.diff(tt[3, 20:220]); print("pace per tile (G handed -> G handed): mean %.0f  p10 %.0f p90 %.0f" % (dj.mean(), np.percentile(dj, 10), np.percentile(dj, 90)))
    print("  epilogue waits for S      : mean %.0f" % (tt[1, 20:220] - tt[0, 20:220]).mean())
    print("  S seen -> exps done       : mean %.0f" % (tt[2, 20:220] - tt[1, 20:220]).mean())
    print("  exps done -> G handed     : mean %.0f" % (tt[3, 20:220] - tt[2, 20:220]).mean())
    print("  G handed -> mma has G     : mean %.0f" % (tt[5, 20:220] - tt[3, 20:220]).mean())
    print("  mma waits for G           : mean %.0f" % (tt[5, 20:220] - tt[4, 20:220]).mean())
    print("  mma has G -> queued       : mean %.0f" % (tt[6, 20:220] - tt[5, 20:220]).mean())
    print("  G(j) handed -> S(j+2) seen: mean %.0f" % (tt[1, 22:222] - tt[3, 20:220]).mean())
    w = tt[8:16, 20:220]
    if (w != 0).all():
        print("  hand-over of epilogue warp e (warp e + 2, sub-partition (e + 2) % 4) relative to the first one: "
              + " ".join("%d" % v for v in (w - w.min(axis=0)).mean(axis=1)))
        print("  last hand-over -> mma has G: mean %.0f" % (tt[5, 20:220] - w.max(axis=0)).mean())
