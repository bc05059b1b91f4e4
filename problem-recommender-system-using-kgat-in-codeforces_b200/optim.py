"""``torch.optim.Adam``-compatible optimiser backed by the multi-tensor Adam kernel (K11).

The reference builds two ``torch.optim.Adam(self.parameters(), lr)`` with default hyper-parameters
(model.py:404-405) and calls ``step()`` / ``zero_grad()`` (model.py:407-419).  ``FusedAdam`` keeps
those semantics -- parameters without a gradient are skipped, state is created lazily per
parameter, each parameter has its own step count -- but updates every tensor of a step in one
kernel launch instead of ~10 ``_foreach`` launches.
"""

from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._hyper: dict[int, torch.Tensor] = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            by_step: dict[int, list] = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            beta1, beta2 = group["betas"]
            for step, ps in by_step.items():
                dev = ps[0].device
                hyper = self._hyper.get((dev, step % 2))
                if hyper is None:
                    hyper = self._hyper[(dev, step % 2)] = torch.empty(8, dtype=torch.float32, device=dev)
                ops.adam_set_hyper(step, group["lr"], beta1, beta2, group["eps"], hyper)
                ops.adam_apply(
                    [p.data for p in ps],
                    [p.grad.contiguous() for p in ps],
                    [self.state[p]["exp_avg"] for p in ps],
                    [self.state[p]["exp_avg_sq"] for p in ps],
                    hyper,
                )
                for p in ps:  # the kernel wrote through raw pointers: tell autograd the data changed
                    torch.autograd.graph.increment_version(p)
        return loss
