"""NVLink peer-memory mailboxes for the small exchanges of the data-parallel path.

A batch-sharded step needs two kinds of exchange: the pair ``{sum(pixel_sel_loc), sum(sel * row_loss)}`` between the
forward and the backward of the prototype loss (utils/loss.py:558-565 over the global batch), and the per-class sums
``[sets*K, C+1]`` float64 between the class-sum sweep and its finaliser (update_class_center_iter / cal_centroid over the
global batch).  With NCCL each is a collective launch (~20 us of latency plus the Python dispatch) on the critical
path; here they are done INSIDE our own kernels (``slcl_proto_rescale_peer``, the reduce kernel of
``slcl_class_centres_update`` / ``slcl_centroids_fwd``, ``slcl_peer_allreduce_f64``), which store 8-byte
``{epoch | 32-bit payload}`` words straight into every peer's mailbox over NVLink and poll their own.

The mailboxes live in ``torch.distributed._symmetric_memory`` (one process per GPU, one box).  Pass a ``PeerMailbox`` as
the ``group=`` argument of ``mpcl_loss_calc`` / ``mpcl_target_step`` / ``update_class_center_iter`` / ``cal_centroid``.
Rules: every rank makes the same sequence of ``group=mailbox`` calls, all on one CUDA stream per mailbox.  When
symmetric memory cannot be set up, pass ``group=True`` (or a ProcessGroup) and the same calls use NCCL all-reduces.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from . import _lib

DEFAULT_CAPACITY_WORDS = 8192        # 4096 doubles per message: [16, 255+1] class sums
DEFAULT_TIMEOUT_S = 600.0            # a rank that checkpoints / evaluates for minutes must not poison the step


class PeerMailbox:
    """One mailbox per rank of ``group``.  Construction is a collective call.

    ``capacity_words``: payload words per message (two per float64).  ``timeout_s``: how long a kernel waits for a
    peer before it gives up and returns NaN (0 = for ever, like an NCCL collective); ``timeouts()`` reads the count of
    such events and ``check()`` raises if there were any."""

    def __init__(self, device, group=None, capacity_words: int = DEFAULT_CAPACITY_WORDS, timeout_s: float = DEFAULT_TIMEOUT_S):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("PeerMailbox needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.dev = torch.device(device)
        pg = dist.group.WORLD if group is None or group is True else group
        self.group = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.capacity_words = int(capacity_words)
        self.timeout_s = float(timeout_s)
        n_bytes = self.lib.slcl_peer_mailbox_bytes(self.world, self.capacity_words)
        if n_bytes == 0:
            raise ValueError(f"peer exchange supports 1..16 ranks and capacity >= 2 words, got world={self.world}")
        self.buf = symm.empty(n_bytes // 8, dtype=torch.int64, device=self.dev)
        self.buf.zero_()
        torch.cuda.synchronize(self.dev)
        self.handle = symm.rendezvous(self.buf, pg)
        self.ptrs_dev = int(self.handle.buffer_ptrs_dev)
        dist.barrier(pg)                      # every mailbox is zeroed and mapped before the first store can arrive
        torch.cuda.synchronize(self.dev)

    def struct(self) -> "_lib.PeerT":
        return _lib.PeerT(self.ptrs_dev, self.rank, self.world, self.capacity_words, self.timeout_s)

    def args(self):
        """(ptrs_dev, rank, world, capacity_words, timeout_s): the flat form the torch custom ops take."""
        return self.ptrs_dev, self.rank, self.world, self.capacity_words, self.timeout_s

    def epoch(self) -> int:
        return int(self.buf[0].item())

    def timeouts(self) -> int:
        return int(self.buf[1].item())

    def check(self) -> None:
        """Host-side check (synchronises): raise if any exchange kernel gave up waiting for a peer."""
        n = self.timeouts()
        if n:
            raise RuntimeError(f"slcl peer exchange: {n} poll(s) timed out after {self.timeout_s} s on rank {self.rank}; "
                               "the affected results are NaN")


class _LoopbackBox(PeerMailbox):
    def __init__(self, owner, rank):        # noqa: D401 - built by LoopbackMailboxes only
        self._owner = owner                 # keeps the pointer table (and every peer's mailbox) alive
        self.lib = owner.lib
        self.dev = owner.dev
        self.group = None
        self.world = owner.world
        self.rank = rank
        self.capacity_words = owner.capacity_words
        self.timeout_s = owner.timeout_s
        self.buf = owner.bufs[rank]
        self.ptrs_dev = owner.table.data_ptr()


class LoopbackMailboxes:
    """``world`` mailboxes in ordinary memory of ONE device (no process group): ``boxes[r]`` is what rank r would hold.
    Several "ranks" can then run on different CUDA streams of one GPU and really exchange through the protocol --
    how the single-GPU test box exercises the multi-rank kernels."""

    def __init__(self, world: int, device, capacity_words: int = DEFAULT_CAPACITY_WORDS, timeout_s: float = 20.0):
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.world = int(world)
        self.capacity_words = int(capacity_words)
        self.timeout_s = float(timeout_s)
        n_bytes = self.lib.slcl_peer_mailbox_bytes(self.world, self.capacity_words)
        if n_bytes == 0:
            raise ValueError("peer exchange supports 1..16 ranks and capacity >= 2 words")
        self.bufs: List[torch.Tensor] = [torch.zeros(n_bytes // 8, dtype=torch.int64, device=self.dev) for _ in range(world)]
        self.table = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device=self.dev)
        torch.cuda.synchronize(self.dev)
        self.boxes = [_LoopbackBox(self, r) for r in range(world)]
