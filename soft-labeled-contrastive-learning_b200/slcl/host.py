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
        if min(b, c, h, w) < 1:
            raise ValueError("feature map must be a non-empty [B, C, h, w]")
        self.shape = tuple(shape)
        self.dev = torch.device(device)
        self.chunk = max(1, min(chunk_images, b))
        self.n_slots = n_slots
        self.with_sel = with_sel
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.centres = torch.empty(n_class, c, **f32)
        self._mk = lambda n_img: self._make_slot(n_img, c, h, w, n_class, mpcl, f32)
        self.slots = [self._mk(self.chunk) for _ in range(n_slots)]
        tail = b % self.chunk                       # a short last chunk gets its own plan (any batch size is accepted,
        self.tail = self._mk(tail) if tail else None                                    # like the reference path)
        self.total = torch.zeros(1, **f32)          # global weight sum (sum(sel) or N)
        self.loss = torch.zeros((), **f32)

    def _make_slot(self, n_img, c, h, w, n_class, mpcl, f32):
        px = n_img * h * w
        feat = torch.empty(n_img, c, h, w, **f32)
        labels = torch.empty(px, dtype=torch.int64, device=self.dev)
        sel = torch.empty(px, **f32) if self.with_sel else None
        plan = ProtoPlan(feat, labels, sel, self.centres, n_class, mpcl.temperature, mpcl.base_temperature, mpcl.m,
                         mpcl.easy_margin, True)
        return dict(feat=feat, labels=labels, sel=sel, plan=plan, stream=torch.cuda.Stream(self.dev),
                    done=torch.cuda.Event(), loss=torch.zeros((), **f32), n_img=n_img)

    def run(self, feas_h: torch.Tensor, labels_h: torch.Tensor, centres: torch.Tensor, sel_h: Optional[torch.Tensor],
            grad_h: torch.Tensor, group=None) -> torch.Tensor:
        # every check comes BEFORE the first collective: a rank that raises must not leave its peers in an all-reduce
        b, c, h, w = self.shape
        hw = h * w
        if tuple(feas_h.shape) != self.shape or tuple(grad_h.shape) != self.shape:
            raise ValueError("feas_h / grad_h do not have the shape this pipeline was built for")
        labels_h = labels_h.reshape(-1)
        if labels_h.numel() != b * hw or labels_h.dtype != torch.int64:
            raise ValueError("labels must be int64 at feature resolution ([B,h,w] or [B*h*w])")
        if self.with_sel != (sel_h is not None) or (sel_h is not None and sel_h.numel() != b * hw):
            raise ValueError("pixel_sel_loc must have B*h*w elements (and match how the pipeline was built)")
        if tuple(centres.shape) != tuple(self.centres.shape):
            raise ValueError("class centres must be [K, C]")
        multi = False
        if group is not None:
            import torch.distributed as dist
            pg = None if group is True else group
            multi = dist.is_initialized() and dist.get_world_size(pg) > 1
        main = torch.cuda.current_stream(self.dev)
        self.centres.copy_(centres, non_blocking=True)
        if self.with_sel:
            self.total.copy_(sel_h.reshape(-1).sum().reshape(1), non_blocking=True)      # host-side sum of a host tensor
        else:
            self.total.fill_(float(b * hw))
        if multi:
            dist.all_reduce(self.total, group=pg)             # global normaliser: sum(sel) (or N) over all ranks
        for sl in self.slots + ([self.tail] if self.tail else []):
            sl["loss"].zero_()
        ready = torch.cuda.Event()
        ready.record(main)
        n_full = b // self.chunk
        work = [(self.slots[i % self.n_slots], i * self.chunk, (i + 1) * self.chunk) for i in range(n_full)]
        if self.tail is not None:
            work.append((self.tail, n_full * self.chunk, b))
        used = []
        for sl, lo, hi in work:
            st = sl["stream"]
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
            if not any(sl is u for u in used):
                used.append(sl)
        for sl in used:
            main.wait_event(sl["done"])
        self.loss.copy_(torch.stack([sl["loss"] for sl in used]).sum())
        if multi:
            # every chunk's share was taken with the GLOBAL normaliser, so the global loss (utils/loss.py:565 over the
            # whole batch) is the plain sum of the ranks' shares; gradients already carry the global coefficient
            dist.all_reduce(self.loss, group=pg)
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
