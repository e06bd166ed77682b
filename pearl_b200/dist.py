"""Multi-GPU data-parallel learner plumbing (one process per GPU).

`torch.distributed` is used only for the rendezvous (rank/world, a byte
all-gather of CUDA IPC handles).  The gradient exchange itself happens inside
the persistent learner kernel over NVLink peer memory (include/pearl_b200.h,
prl_comm_*): the reference has no distributed RL learner, this
is the north star's "all-reduce on the gradient only".
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib

HANDLE_BYTES = 128


def all_gather_bytes(blob: bytes, group=None, device: Optional[torch.device] = None) -> List[bytes]:
    """All-gather one fixed-size byte string per rank (works with nccl and gloo)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = device if (backend == "nccl" and device is not None) else (
        torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().tolist()) for t in out]


def shard_owner(global_write_index: int, world: int) -> tuple:
    """Interleaved ownership of a globally numbered transition stream:
    transition g lives on rank g mod W at local position g div W, which keeps FIFO
    eviction and age-uniform sampling balanced across shards."""
    return global_write_index % world, global_write_index // world


def sharded_gae_fixup(v0: float, g0: float, has_cut: bool, redo, group=None, device: Optional[torch.device] = None) -> int:
    """GAE over a rollout sharded by contiguous time chunks, rank r holding chunk r (the recurrence of
    ppo.py:271-293 runs newest -> oldest, so chunk r continues into chunk r + 1).

    After its local pass every rank knows `v0` (state value of its oldest transition), `g0` (gae of its oldest
    transition) and `has_cut` (the chunk contains a terminated / truncated transition, which makes g0 independent of
    newer chunks).  One all-gather of 4 floats per rank and round; a rank whose successor's g0 is final calls
    `redo(next_value, incoming_gae)` once (it recomputes the chunk's chains and returns the new g0).  Chunks without an
    episode end propagate sequentially, one round per such chunk; otherwise two rounds.  Returns the number of rounds."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    final = bool(has_cut) or rank == world - 1     # g0 no longer depends on newer chunks
    done = rank == world - 1                       # every element of the chunk is final
    rounds = 0
    while True:
        mine = torch.tensor([v0, g0, float(final), float(done)], dtype=torch.float32, device=dev)
        table = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(table, mine, group=group)
        table = [t.cpu() for t in table]
        rounds += 1
        if all(bool(t[3]) for t in table):
            return rounds
        if not done and bool(table[rank + 1][2]):
            g0 = float(redo(float(table[rank + 1][0]), float(table[rank + 1][1])))
            final = done = True
        if rounds > world + 1:
            raise RuntimeError("sharded GAE did not converge (inconsistent shards?)")


class B200Communicator:
    """NVLink peer-memory communicator for B200DeepQLearning / B200DoubleDQN."""

    def __init__(self, param_count: int, device: torch.device, group=None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self._lib = _lib.init(self.device.index)
        self._handle = C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.prl_comm_create(C.byref(self._handle), self.rank, self.world, int(param_count)))
            blob = (C.c_uint8 * HANDLE_BYTES)()
            _lib.check(self._lib.prl_comm_local_handles(self._handle, blob))
            blobs = all_gather_bytes(bytes(blob), group, self.device)
            joined = (C.c_uint8 * (HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(blobs))
            _lib.check(self._lib.prl_comm_open_peers(self._handle, joined))
        dist.barrier(group)

    @property
    def handle(self) -> C.c_void_p:
        return self._handle

    def close(self) -> None:
        if self._handle.value:
            self._lib.prl_comm_destroy(self._handle)
            self._handle = C.c_void_p(0)
