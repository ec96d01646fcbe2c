"""Time rp_peer_allreduce alone (torchrun, one process per GPU): python -m torch.distributed.run --nproc-per-node N tools/time_peer_allreduce.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
from replay_b200.peer import alloc_peer_grad
n = 6_625_152
peer = alloc_peer_grad(n, torch.device("cuda", local))
st = torch.cuda.current_stream().cuda_stream
ref = torch.randn(n, device="cuda")
def t(fn, it=30):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
tp = t(lambda: peer.all_reduce(st)) if peer is not None else float("nan")
tn = t(lambda: dist.all_reduce(ref))
if rank == 0:
    print(f"world {world}: rp_peer_allreduce {tp * 1e3:.1f} us, ncclAllReduce {tn * 1e3:.1f} us for {n * 4 / 1e6:.1f} MB  (lib {os.environ.get('RP_B200_LIB', 'default')})")
dist.destroy_process_group()
