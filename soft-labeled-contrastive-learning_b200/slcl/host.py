"""Host-buffer entry point of the prototype loss: features / labels / selection mask live in
(pinned) HOST memory, the gradient comes back to a host buffer.

The batch is cut into image chunks that flow through a small ring of device buffers on separate
CUDA streams, so the host->device copy of chunk i+1, the kernels of chunk i and the device->host
copy of the gradient of chunk i-1 overlap (PCIe is full duplex).  The loss normaliser
(sum of ``pixel_sel_loc``, utils/loss.py:565) is known before the first chunk, so every chunk's
backward can run right behind its forward with the GLOBAL coefficient.

Same arithmetic as ``mpcl_loss_calc(...)`` + ``.backward()`` (reference utils/loss.py:576-605,
484-573); the kernels are the ones of the device path (slcl.plan.ProtoPlan).
"""
from __future__ import annotations

from typing import Optional

import torch

from .loss import MPCL
from .plan import ProtoPlan


class HostProtoPipeline:
    def __init__(self, shape, n_class: int, mpcl: MPCL, with_sel: bool, device, chunk_images: int = 4, n_slots: int = 3):
        b, c, h, w = shape
        self.shape = tuple(shape)
        self.dev = torch.device(device)
        self.chunk = max(1, min(chunk_images, b))
        self.n_slots = n_slots
        self.with_sel = with_sel
        px = self.chunk * h * w
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.centres = torch.empty(n_class, c, **f32)
        self.slots = []
        for _ in range(n_slots):
            feat = torch.empty(self.chunk, c, h, w, **f32)
            labels = torch.empty(px, dtype=torch.int64, device=self.dev)
            sel = torch.empty(px, **f32) if with_sel else None
            plan = ProtoPlan(feat, labels, sel, self.centres, n_class, mpcl.temperature, mpcl.base_temperature, mpcl.m,
                             mpcl.easy_margin, True)
            self.slots.append(dict(feat=feat, labels=labels, sel=sel, plan=plan, stream=torch.cuda.Stream(self.dev),
                                   done=torch.cuda.Event(), loss=torch.zeros((), **f32)))
        self.total = torch.zeros(1, **f32)          # global weight sum (sum(sel) or N)
        self.loss = torch.zeros((), **f32)

    def run(self, feas_h: torch.Tensor, labels_h: torch.Tensor, centres: torch.Tensor, sel_h: Optional[torch.Tensor],
            grad_h: torch.Tensor, group=None) -> torch.Tensor:
        b, c, h, w = self.shape
        hw = h * w
        main = torch.cuda.current_stream(self.dev)
        labels_h = labels_h.reshape(-1)
        self.centres.copy_(centres, non_blocking=True)
        if self.with_sel:
            self.total.copy_(sel_h.reshape(-1).sum().reshape(1), non_blocking=True)      # host-side sum of a host tensor
        else:
            self.total.fill_(float(b * hw))
        if group is not None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(None if group is True else group) > 1:
                dist.all_reduce(self.total, group=None if group is True else group)
        for sl in self.slots:
            sl["loss"].zero_()
        ready = torch.cuda.Event()
        ready.record(main)
        n_chunks = (b + self.chunk - 1) // self.chunk
        if b % self.chunk:
            raise ValueError("batch must be a multiple of chunk_images")
        for i in range(n_chunks):
            sl = self.slots[i % self.n_slots]
            st = sl["stream"]
            lo, hi = i * self.chunk, (i + 1) * self.chunk
            with torch.cuda.stream(st):
                st.wait_event(ready)
                sl["feat"].copy_(feas_h[lo:hi], non_blocking=True)
                sl["labels"].copy_(labels_h[lo * hw:hi * hw], non_blocking=True)
                if self.with_sel:
                    sl["sel"].copy_(sel_h.reshape(-1)[lo * hw:hi * hw], non_blocking=True)
                plan = sl["plan"]
                scal = plan.forward()
                scal[2:3].copy_(self.total)          # global normaliser -> scal[0] = this chunk's share of the loss
                plan.rescale()
                dfeat = plan.backward()
                grad_h[lo:hi].copy_(dfeat, non_blocking=True)
                sl["loss"].add_(scal[0])             # per-slot accumulator: streams never share a destination
                sl["done"].record(st)
        for sl in self.slots:
            main.wait_event(sl["done"])
        self.loss.copy_(torch.stack([sl["loss"] for sl in self.slots]).sum())
        return self.loss


_cache = {}


def mpcl_loss_and_grad_host(feas_h: torch.Tensor, labels_h: torch.Tensor, class_center_feas: torch.Tensor, loss_func: MPCL,
                            pixel_sel_loc_h: Optional[torch.Tensor] = None, grad_out_h: Optional[torch.Tensor] = None,
                            device="cuda", chunk_images: int = 4, group=None):
    """feas_h [B,C,h,w] fp32 and labels / pixel_sel_loc in host memory (pin them for full speed);
    returns (loss: 0-d device tensor, grad_h: host tensor with dloss/dfeas).  Labels must be at feature
    resolution ([B,h,w] or [B*h*w]).  Asynchronous w.r.t. the host: synchronise (or read ``loss.item()``)
    before touching ``grad_h``."""
    if not isinstance(loss_func, MPCL):
        raise TypeError("loss_func must be an slcl.loss.MPCL")
    if feas_h.is_cuda:
        raise ValueError("feas_h must be a host tensor; use slcl.loss.mpcl_loss_calc for device tensors")
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if grad_out_h is None:
        grad_out_h = torch.empty_like(feas_h, pin_memory=True)
    key = (tuple(feas_h.shape), loss_func.num_class, pixel_sel_loc_h is not None, dev, chunk_images,
           loss_func.temperature, loss_func.base_temperature, loss_func.m, loss_func.easy_margin)
    pipe = _cache.get(key)
    if pipe is None:
        _cache.clear()
        pipe = HostProtoPipeline(feas_h.shape, loss_func.num_class, loss_func, pixel_sel_loc_h is not None, dev, chunk_images)
        _cache[key] = pipe
    loss = pipe.run(feas_h, labels_h, class_center_feas, pixel_sel_loc_h, grad_out_h, group)
    return loss, grad_out_h
