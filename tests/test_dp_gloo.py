"""world_size-2 `gloo` test (CPU) of the data-parallel host logic: gradient all-reduce + 1/world scaling in Trainer.step and the
exact contiguous user sharding of predict()."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from replay_b200.trainer import user_shard


class _FakeLib:
    count = 0


class _FakeCE:
    loss = torch.zeros(2)


class FakeEngine:
    """Stands in for SasRecEngine: 'backward' writes a rank-dependent gradient; 'optimizer_step' records what it was given."""

    def __init__(self, rank):
        self.rank = rank
        self.g32 = torch.zeros(8)
        self.lib, self.ce = _FakeLib(), _FakeCE()
        self.seen_scale, self.seen_grad = None, None

    def set_batch(self, *a):
        self.batch = a

    def tick_rng(self):
        pass

    def forward_train(self):
        return self.ce.loss

    def backward(self):
        self.g32 = torch.arange(8, dtype=torch.float32) * (self.rank + 1)

    def optimizer_step(self, grad_scale=1.0, **kw):
        self.seen_scale, self.seen_grad = grad_scale, self.g32.clone()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from replay_b200.trainer import Trainer

    eng = FakeEngine(rank)
    tr = Trainer(eng, use_graph=False)
    assert tr.world == world
    tr.step(None, None, None, None)
    # the callback the Lightning mirrors' fused training_step hands to the engine (replay_b200/core.py): sum + 1/world
    from replay_b200.core import dist_grad_all_reduce

    cb = dist_grad_all_reduce()
    g = torch.arange(4, dtype=torch.float32) + 10 * rank
    scale = cb(g)
    assert scale == 0.5 and g.tolist() == [10.0, 12.0, 14.0, 16.0], (scale, g)
    q.put((rank, eng.seen_scale, eng.seen_grad.tolist()))
    dist.destroy_process_group()


def test_trainer_all_reduce_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [float(i * 3) for i in range(8)]  # rank0 grad + rank1 grad = i*1 + i*2
    for rank, scale, grad in res:
        assert scale == 0.5
        assert grad == expect


def test_user_shard_is_exact_partition():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            cuts = [user_shard(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_grad_all_reduce_callback_is_none_without_a_process_group():
    from replay_b200.core import dist_grad_all_reduce

    assert dist_grad_all_reduce() is None
