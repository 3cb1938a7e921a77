"""NVLink peer-memory mailboxes for the scalar exchanges of the data-parallel path.

The prototype loss of a batch-sharded step needs ONE exchange between forward and backward: the pair
``{sum(pixel_sel_loc), sum(sel * row_loss)}`` (utils/loss.py:558-565 over the global batch).  With NCCL that is
a collective launch (~20 us of latency) on the critical path; ``slcl_proto_rescale_peer`` does the exchange and
the rescale in ONE kernel that stores 8-byte ``{epoch | fp32}`` words straight into every peer's mailbox over
NVLink and polls its own.  The mailboxes live in ``torch.distributed._symmetric_memory`` (one process per GPU,
one box).  When symmetric memory cannot be set up the callers keep using the NCCL all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class PeerMailbox:
    """One mailbox per rank of ``group``; ``ptrs_dev`` is the device array of peer mailbox pointers that
    ``slcl_proto_rescale_peer`` takes.  Construction is a collective call."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("PeerMailbox needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.dev = torch.device(device)
        pg = dist.group.WORLD if group is None or group is True else group
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        n_bytes = self.lib.slcl_peer_mailbox_bytes(self.world)
        if n_bytes == 0:
            raise ValueError(f"peer exchange supports 1..16 ranks, got {self.world}")
        self.buf = symm.empty(n_bytes // 8, dtype=torch.int64, device=self.dev)
        self.buf.zero_()
        torch.cuda.synchronize(self.dev)
        self.handle = symm.rendezvous(self.buf, pg)
        self.ptrs_dev = int(self.handle.buffer_ptrs_dev)
        dist.barrier(pg)                      # every mailbox is zeroed and mapped before the first store can arrive
        torch.cuda.synchronize(self.dev)

    def timeouts(self) -> int:
        return int(self.buf[1].item())
