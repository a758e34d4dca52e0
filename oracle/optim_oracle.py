"""TEST INFRASTRUCTURE ONLY - the reference's optimiser step replayed with torch's own CPU classes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The reference's step IS a sequence of torch calls (focoos/trainer/trainer.py:757-773, focoos/trainer/solver/build.py:29-37,104-138):
    grad_scaler.scale(losses).backward(); grad_scaler.unscale_(opt); clip_grad_norm_(params, 0.1)
    grad_scaler.step(opt)  ->  FullModelGradientClippingOptimizer.step: clip_grad_norm_(all_params, 0.1); AdamW.step
    grad_scaler.update()
so the oracle is those calls, on CPU tensors (torch.amp.GradScaler("cpu", init_scale=2**10), torch.optim.AdamW(foreach=False)).
"""
from __future__ import annotations

import itertools

import torch


class ReferenceStepper:
    def __init__(self, param_groups, lr, weight_decay, clip=0.1, amp=True, init_scale=2.0 ** 10, growth_interval=2000):
        class _Clipping(torch.optim.AdamW):  # build.py:29-37
            def step(self, closure=None):
                all_params = itertools.chain(*[x["params"] for x in self.param_groups])
                torch.nn.utils.clip_grad_norm_(all_params, clip)
                super().step(closure=closure)

        groups = [{k: v for k, v in g.items() if k != "name"} for g in param_groups]
        self.opt = (_Clipping if clip > 0 else torch.optim.AdamW)(groups, lr=lr, weight_decay=weight_decay, foreach=False)
        self.scaler = torch.amp.GradScaler("cpu", init_scale=init_scale, growth_interval=growth_interval, enabled=amp)
        self.clip = clip
        self.params = [g["params"][0] for g in param_groups]

    def step(self, loss_fn, world_grads=None):
        """loss_fn() -> scalar loss.  world_grads: optional callable applied after backward (e.g. to average over ranks)."""
        self.opt.zero_grad()
        loss = loss_fn()
        self.scaler.scale(loss).backward()
        if world_grads is not None:
            world_grads(self.params)
        self.scaler.unscale_(self.opt)
        if self.clip > 0:
            ps = [p for p in self.params if p.requires_grad and p.grad is not None]
            torch.nn.utils.clip_grad_norm_(ps, self.clip)
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss.detach()
