"""Fused Adam + finite-gradient guard behind the reference's optimizer objects.

The reference builds `optim.Adam(model.parameters(), lr=...)` (train/loop_utils.py:46-52), guards every step with a
per-parameter NaN/Inf Python loop (train/multimodal.py:141-145) and lets StepLR rewrite `param_groups[0]["lr"]` per epoch.
`FusedAdam.adopt(optimizer, flat)` keeps that torch.optim.Adam object as the user-facing handle (schedulers, lr logging and
`state_dict()` keep working) but re-homes parameters, exp_avg and exp_avg_sq into flat fp32 buffers laid out like the flat
gradient buffer (mauv.flatgrad.FlatGrads), so that one C-ABI call - mauv_adam_step_f32: finite check, step bookkeeping,
update; three launches, 32 B per parameter - replaces the guard loop and the 696-tensor multi-tensor update. The decision to
skip a step with non-finite gradients is taken on the device. There is no PyTorch fallback inside: `adopt` returns None when
the optimizer cannot be expressed (other class, amsgrad, several groups, foreign parameters) and the caller keeps using
`optimizer.step()` itself.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from .flatgrad import FlatGrads


class FusedAdam:
    def __init__(self, optimizer: torch.optim.Adam, flat: FlatGrads):
        self.optimizer, self.flat = optimizer, flat
        self.group = optimizer.param_groups[0]
        dev = flat.flat.device
        n = flat.flat.numel()
        self.p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        lib = _lib.require_device()
        self.state = torch.zeros(lib.mauv_adam_state_bytes() // 4, dtype=torch.int32, device=dev)
        base = flat.flat.data_ptr()
        steps = set()
        for prm, gview in flat.views:
            o = (gview.data_ptr() - base) // 4
            k = prm.numel()
            pv = self.p[o:o + k].view_as(prm)
            pv.copy_(prm.data)
            prm.data = pv                                   # same Parameter object, storage inside the flat buffer
            st = optimizer.state.get(prm, {})
            mv, vv = self.m[o:o + k].view_as(prm), self.v[o:o + k].view_as(prm)
            if "exp_avg" in st:                             # adopt an optimizer that has already stepped
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                steps.add(int(st["step"]))
            # the torch object's state points at our buffers, so optimizer.state_dict() stays meaningful
            optimizer.state[prm] = {"step": torch.tensor(0.0), "exp_avg": mv, "exp_avg_sq": vv}
        if len(steps) > 1:
            raise _lib.MauvError("FusedAdam: parameters with different step counts cannot share one fused update")
        if steps:
            self.state[0] = steps.pop()

    @staticmethod
    def adopt(optimizer, flat: Optional[FlatGrads]) -> Optional["FusedAdam"]:
        """FusedAdam for a plain torch.optim.Adam whose single parameter group is exactly the flat buffer's parameters."""
        if flat is None or type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
            return None
        g = optimizer.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable"):
            return None
        if isinstance(g["lr"], torch.Tensor) or not flat.flat.is_cuda:
            return None
        ours = [p for p, _ in flat.views]
        theirs = [p for p in g["params"] if p.requires_grad]
        if len(ours) != len(theirs) or any(a is not b for a, b in zip(ours, theirs)):
            return None
        return FusedAdam(optimizer, flat)

    def step(self) -> torch.Tensor:
        """One guarded Adam step on the device. -> 0-d int32 tensor view: 1 iff the update was applied (no sync here)."""
        g = self.group
        b1, b2 = g["betas"]
        self.flat.ensure_attached()
        ops.adam_step_f32(self.p, self.flat.flat, self.m, self.v, float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                          float(g.get("weight_decay", 0.0)), self.state)
        self.optimizer._opt_called = True           # lr_scheduler.step() checks that the optimizer has stepped
        return self.state[1]

    @property
    def applied(self) -> torch.Tensor:
        return self.state[1]

    def sync_state(self) -> int:
        """Copy the device step count into the torch optimizer's per-parameter state (for optimizer.state_dict())."""
        t = int(self.state[0])
        for prm, _ in self.flat.views:
            self.optimizer.state[prm]["step"] = torch.tensor(float(t))
        return t
