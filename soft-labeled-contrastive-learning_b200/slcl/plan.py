"""Low-overhead launch plans over the C ABI.

A plan owns every output / stash / workspace buffer of one fixed-shape call and
pre-builds the ctypes argument list, so a step is a couple of raw C calls (a few
microseconds of host time) instead of the torch custom-op dispatch path.  The
launches are stream-ordered and allocation-free, so ``capture_graph()`` can
record forward + backward into one CUDA graph for launch-bound shapes (cfg1:
8 712 pixels).  The trainers' drop-in API (slcl.loss / slcl.utils_) does not need
this; it exists for callers that drive the kernels every step on static shapes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import MapT, ProtoParamsT, check, ptr
from .ops import _map_nchw


class ProtoPlan:
    """Prototype-loss forward/backward (reference mpcl_loss_calc + MPCL.forward,
    utils/loss.py:576-605,484-573) on fixed device buffers."""

    def __init__(self, feat: torch.Tensor, labels: torch.Tensor, sel: Optional[torch.Tensor], centres: torch.Tensor,
                 n_class: int, temperature: float, base_temperature: float, margin: float, easy_margin: bool = False,
                 normalize: bool = True):
        self.lib = _lib.load()
        self.dev = _lib.require_cuda(feat, labels, sel, centres)
        self.feat, self.map = _map_nchw(feat)
        if self.feat.data_ptr() != feat.data_ptr():
            raise ValueError("ProtoPlan needs a feature map whose H,W collapse to one stride (e.g. contiguous NCHW)")
        n = self.map.batch * self.map.pixels
        self.n_pixels = n
        self.labels = labels.reshape(-1).contiguous()
        self.sel = None if sel is None else sel.reshape(-1).float().contiguous()
        self.centres = centres.float().contiguous()
        if self.labels.dtype != torch.int64 or self.labels.numel() != n:
            raise ValueError("labels must be int64 with B*H*W elements")
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.params = ProtoParamsT(int(n_class), float(temperature), float(base_temperature), float(margin),
                                   int(easy_margin), int(normalize))
        self.scal = torch.empty(4, **f32)
        self.stash = torch.empty((n_class + 1, n), **f32)
        self.cstate = torch.empty(n_class * self.map.channels + n_class, **f32)
        self.ws = torch.empty(max(self.lib.slcl_proto_workspace_bytes(n), 256), dtype=torch.uint8, device=self.dev)
        self.dfeat = torch.empty_strided(self.feat.shape, self.feat.stride(), **f32)
        self.grad_out = torch.ones(1, **f32)
        self._fwd_args = (ptr(self.feat), C.byref(self.map), ptr(self.labels), None, ptr(self.sel), ptr(self.centres),
                          C.byref(self.params), ptr(self.stash), ptr(self.cstate), ptr(self.scal), ptr(self.ws),
                          self.ws.numel())
        self._bwd_args = (ptr(self.feat), C.byref(self.map), ptr(self.stash), ptr(self.cstate), ptr(self.scal),
                          ptr(self.grad_out), C.byref(self.params), ptr(self.dfeat))
        self.graph = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def forward(self, mailbox=None, split_phase: bool = False) -> torch.Tensor:
        """Launch the forward; returns scal (scal[0] is the loss).  Asynchronous.  With ``mailbox`` (slcl.peer.PeerMailbox,
        more than one rank) the finaliser kernel exchanges the loss pair with the other ranks: scal is then GLOBAL and
        neither ``rescale()`` nor ``rescale_peer()`` is needed before ``backward()``.

        ``split_phase=True``: the finaliser only SENDS; ``backward(mailbox)`` -- which MUST be the next use of the mailbox
        -- receives, so the exchange latency hides behind the backward's launch; scal is global after that backward."""
        if mailbox is not None and mailbox.world > 1:
            peer = mailbox.struct()
            args = self._fwd_args[:10] + (C.byref(peer), int(split_phase)) + self._fwd_args[10:]
            check(self.lib.slcl_proto_fwd_peer(*args, self._stream()), "slcl_proto_fwd_peer")
        else:
            check(self.lib.slcl_proto_fwd(*self._fwd_args, self._stream()), "slcl_proto_fwd")
        return self.scal

    def rescale(self) -> None:
        check(self.lib.slcl_proto_rescale(ptr(self.scal), int(self.sel is not None), self._stream()), "slcl_proto_rescale")

    def rescale_peer(self, mailbox) -> None:
        """Exchange {weight sum, weighted row-loss sum} with the other ranks through NVLink peer memory and rescale,
        in ONE kernel (``mailbox``: slcl.peer.PeerMailbox); replaces all_reduce(scal[2:4]) + rescale()."""
        peer = mailbox.struct()
        check(self.lib.slcl_proto_rescale_peer(ptr(self.scal), int(self.sel is not None), C.byref(peer), self._stream()),
              "slcl_proto_rescale_peer")

    def backward(self, mailbox=None) -> torch.Tensor:
        """Launch the backward for dL/dloss = grad_out (device scalar, default 1); returns dfeat.  ``mailbox``: completes a
        ``forward(mailbox, split_phase=True)``."""
        if mailbox is not None and mailbox.world > 1:
            peer = mailbox.struct()
            check(self.lib.slcl_proto_bwd_peer(*self._bwd_args, C.byref(peer), int(self.sel is not None), self._stream()),
                  "slcl_proto_bwd_peer")
        else:
            check(self.lib.slcl_proto_bwd(*self._bwd_args, self._stream()), "slcl_proto_bwd")
        return self.dfeat

    def capture_graph(self) -> "torch.cuda.CUDAGraph":
        """Record forward + backward into one CUDA graph (replay with ``plan.graph.replay()``)."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.forward(); self.backward()          # warm-up outside capture
        torch.cuda.current_stream(self.dev).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward()
            self.backward()
        self.graph = g
        return g


class P2PPlan:
    """Pixel<->pixel loss forward + backward (slcl_p2p_fwd / slcl_p2p_bwd) on fixed device buffers:
    anchors a [A, dp] and contrast rows b [M, dp] (bf16, dp % 64 == 0), padded int32 metadata
    (slcl.ops.pad_meta), shift / weight [A].  ``n_class`` > 0 selects the analytic sweeps (class-index labels,
    self-pair maps a_selfcol / b_selfrow).  Outputs: loss[1], stats[A,3], d_a [A, dim], d_b [M, dim]."""

    def __init__(self, a: torch.Tensor, b: torch.Tensor, dim: int, a_meta: torch.Tensor, b_meta: torch.Tensor,
                 shift: torch.Tensor, weight: torch.Tensor, temperature: float, n_class: int = 0,
                 a_selfcol: torch.Tensor = None, b_selfrow: torch.Tensor = None):
        self.lib = _lib.load()
        self.dev = _lib.require_cuda(a, b, a_meta, b_meta, shift, weight)
        na, dp = a.shape
        m = b.shape[0]
        self.keep = (a, b, a_meta, b_meta, shift, weight, a_selfcol, b_selfrow)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.stats = torch.empty((na, 3), **f32)
        self.loss = torch.empty(1, **f32)
        self.d_a = torch.empty((na, dim), **f32)
        self.d_b = torch.empty((m, dim), **f32)
        self.grad_out = torch.ones(1, **f32)
        self.state = (torch.empty(self.lib.slcl_p2p_state_bytes(na, dp), dtype=torch.uint8, device=self.dev)
                      if n_class > 0 else None)
        self.ws = torch.empty(max(self.lib.slcl_p2p_workspace_bytes(na, m, dp), 256), dtype=torch.uint8, device=self.dev)
        t = C.c_float(float(temperature))
        self._fwd_args = (ptr(a), ptr(b), na, m, dp, ptr(a_meta), ptr(b_meta), ptr(a_selfcol), int(n_class), 1, ptr(shift),
                          ptr(weight), t, ptr(self.stats), ptr(self.loss), ptr(self.state), ptr(self.ws), self.ws.numel())
        self._fwd_only_args = self._fwd_args[:15] + (None,) + self._fwd_args[16:]
        self._bwd_args = (ptr(a), ptr(b), na, m, dp, dim, ptr(a_meta), ptr(b_meta), ptr(a_selfcol), ptr(b_selfrow),
                          int(n_class), 1, ptr(shift), ptr(weight), t, ptr(self.stats), ptr(self.state), ptr(self.grad_out),
                          ptr(self.d_a), ptr(self.d_b), ptr(self.ws), self.ws.numel())
        self.flops = 8.0 * na * m * dp
        self.graph = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def forward(self) -> torch.Tensor:
        """forward that also keeps what the backward needs (analytic mode: the state buffer)"""
        check(self.lib.slcl_p2p_fwd(*self._fwd_args, self._stream()), "slcl_p2p_fwd")
        return self.loss

    def forward_only(self) -> torch.Tensor:
        """loss and statistics only (no-grad evaluation)"""
        check(self.lib.slcl_p2p_fwd(*self._fwd_only_args, self._stream()), "slcl_p2p_fwd")
        return self.loss

    def backward(self):
        check(self.lib.slcl_p2p_bwd(*self._bwd_args, self._stream()), "slcl_p2p_bwd")
        return self.d_a, self.d_b

    def capture_graph(self) -> "torch.cuda.CUDAGraph":
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.forward(); self.backward()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward()
            self.backward()
        self.graph = g
        return g
