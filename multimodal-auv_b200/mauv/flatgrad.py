"""One contiguous fp32 buffer behind every `.grad` of a model: one memset per step, one finite check, and (data-parallel
training) ONE all-reduce per step instead of one per parameter. Host logic only - torch supplies the memory
and torch.distributed the collective; works on any device/backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch


class FlatGrads:
    ALIGN = 32      # elements: every view starts on a 128-byte boundary (vector loads in the kernels that accumulate into it)

    def __init__(self, params: Iterable[torch.nn.Parameter], device=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGrads: no trainable parameters")
        device = device if device is not None else self.params[0].device
        offs, tot = [], 0
        for p in self.params:
            offs.append(tot)
            tot += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(tot, dtype=torch.float32, device=device)
        self.views: List[Tuple[torch.nn.Parameter, torch.Tensor]] = []
        for p, o in zip(self.params, offs):
            v = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
            self.views.append((p, v))

    def zero(self) -> None:
        """zero every gradient (one memset) and re-attach views an optimizer.zero_grad(set_to_none=True) dropped"""
        self.flat.zero_()
        for p, v in self.views:
            p.grad = v

    def ensure_attached(self) -> None:
        """Before accumulating: parameters whose `.grad` was set to None get their (zeroed) view back."""
        missing = [(p, v) for p, v in self.views if p.grad is None]
        if len(missing) == len(self.views):
            self.flat.zero_()
        for p, v in missing:
            if len(missing) != len(self.views):
                v.zero_()
            p.grad = v

    def all_reduce_mean(self, group=None) -> None:
        """gradients <- mean over the ranks of the process group (equal shard sizes: the global-minibatch mean)"""
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def finite(self) -> torch.Tensor:
        """0-d bool tensor: the reference's per-parameter NaN/Inf guard (train/multimodal.py:141-145) in one pass"""
        return torch.isfinite(self.flat).all()
