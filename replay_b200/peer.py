"""Symmetric (peer-mapped) gradient buffer for the in-graph NVLink all-reduce kernel (csrc/rp_peer_allreduce.cu).

``torch.distributed._symmetric_memory`` is used for the plumbing only: it allocates the buffer with the CUDA virtual-memory
API, exchanges the shareable handles over the process group's store and maps every rank's copy into this process.  The
reduction itself is this repo's kernel.  Falls back (returns ``None``) when there is no NCCL process group with 2..8 ranks on
one node, when ``RP_PEER_ALLREDUCE=0`` or when the symmetric allocation is refused: the trainer then keeps the
``ncclAllReduce`` between two graph replays."""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from ._lib import check, lib


class PeerGrad:
    """The flat fp32 gradient of this rank (``.g32``) inside a symmetric allocation, plus what the kernel needs."""

    def __init__(self, buf, hdl, n, n_pad):
        self.buf, self.hdl, self.n = buf, hdl, n
        self.world, self.rank = hdl.world_size, hdl.rank
        self.g32 = buf[:n]
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        arr = ctypes.c_void_p * self.world
        self._bufs = arr(*ptrs)
        self._states = arr(*[p + n_pad * 4 for p in ptrs])

    def all_reduce(self, stream) -> None:
        """Enqueue the sum-all-reduce of ``g32`` (every rank must do so once per step; CUDA-graph capturable)."""
        check(lib().rp_peer_allreduce(self._bufs, self._states, self.rank, self.world, self.n, stream), "rp_peer_allreduce")


def alloc_peer_grad(n: int, device) -> PeerGrad | None:
    if os.environ.get("RP_PEER_ALLREDUCE", "1") == "0":
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return None
    world = dist.get_world_size()
    if world < 2 or world > 8:
        return None
    peer, err = None, None
    try:
        import torch.distributed._symmetric_memory as symm_mem

        n_pad = (n + 3) // 4 * 4
        state_words = lib().rp_peer_allreduce_state_bytes() // 4
        buf = symm_mem.empty(n_pad + (state_words + 3) // 4 * 4, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        buf.zero_()
        torch.cuda.synchronize(device)
        peer = PeerGrad(buf, hdl, n, n_pad)
    except Exception as e:  # noqa: BLE001 - any refusal (driver, container, torch build) means "use NCCL"
        err = e
    # every rank must take the same path: agree on the outcome (this collective also is the barrier after which nobody's
    # state block is zeroed any more - flags are only raised by the kernel, i.e. later)
    ok = torch.tensor([1 if peer is not None else 0], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        if dist.get_rank() == 0:
            why = f"{type(err).__name__}: {err}" if err is not None else "refused on another rank"
            print(f"[replay_b200] symmetric gradient buffer unavailable ({why}); using ncclAllReduce", flush=True)
        return None
    return peer
