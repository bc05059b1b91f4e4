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
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, use_graphs: bool = True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._hyper: dict[int, torch.Tensor] = {}
        # When the same (parameter, gradient-buffer) set shows up step after step -- the model's API fast path hands
        # autograd static gradient buffers -- the whole update (bias-correction scalars from a device step counter +
        # the multi-tensor kernel) is captured once as a CUDA graph and replayed.
        self.use_graphs = use_graphs
        self._plans: dict = {}
        self._seen: set = set()
        self._fast: dict = {}

    def _graphed_step(self, group, ps) -> bool:
        if not self.use_graphs or not ps or torch.cuda.is_current_stream_capturing():
            return False
        steps = {self.state[p]["step"] if self.state[p] else 0 for p in ps}
        if len(steps) != 1 or not all(p.is_cuda for p in ps):
            return False
        cur = steps.pop()
        beta1, beta2 = group["betas"]
        key = (tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps), group["lr"], beta1, beta2, group["eps"])
        plan = self._plans.get(key)
        if plan is None:
            if key not in self._seen or cur == 0:  # first sighting: take the generic path (also warms the kernels up)
                if len(self._seen) > 64:
                    self._seen.clear()
                self._seen.add(key)
                return False
            if len(self._plans) > 8:
                self._plans.clear()
            dev = ps[0].device
            plan = {"step_dev": torch.full((1,), cur, dtype=torch.int64, device=dev), "hyper": torch.empty(8, dtype=torch.float32, device=dev),
                    "count": cur, "graph": torch.cuda.CUDAGraph()}
            params, grads = [p.data for p in ps], [p.grad for p in ps]
            ms, vs = [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps]
            torch.cuda.synchronize()
            with torch.cuda.graph(plan["graph"]):
                ops.adam_advance(plan["step_dev"], group["lr"], beta1, beta2, group["eps"], plan["hyper"])
                ops.adam_apply(params, grads, ms, vs, plan["hyper"])
            self._plans[key] = plan
        if plan["count"] != cur:  # someone else advanced these parameters: resynchronise the device counter
            plan["step_dev"].fill_(cur)
            plan["count"] = cur
        plan["graph"].replay()
        plan["count"] += 1
        for p in ps:
            self.state[p]["step"] += 1
            torch.autograd.graph.increment_version(p)
        return True

    # -- the API fast path (functions.GraphedStep.try_fused_update) ---------------------------------
    def fast_plan(self, ps, grads, grads_key=None, row_slot0=None):
        """Captured ``adam_advance + adam_apply`` over exactly ``(ps, grads)`` (static gradient buffers of a graphed step),
        or None when it cannot be used yet: first sighting (the generic path creates the state and warms the kernels up),
        parameters with different step counts, graphs disabled, or some other parameter of the group holds a gradient."""
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            return None
        group = self.param_groups[0]
        n_with = 0
        for p in group["params"]:
            if p.grad is not None:
                n_with += 1
        if n_with != len(ps) or len(self.param_groups) != 1:
            return None
        beta1, beta2 = group["betas"]
        if grads_key is None:
            grads_key = tuple(g.data_ptr() for g in grads)
        key = (id(ps[0]), len(ps), grads_key, group["lr"], beta1, beta2, group["eps"])
        plan = self._fast.get(key)
        if plan is None:
            steps = {int(self.state[p]["step"]) if self.state[p] else 0 for p in ps}
            if len(steps) != 1 or 0 in steps:
                return None
            if len(self._fast) > 8:
                self._fast.clear()
            dev = ps[0].device
            cur = steps.pop()
            plan = {"step_dev": torch.full((1,), cur, dtype=torch.int64, device=dev), "hyper": torch.empty(8, dtype=torch.float32, device=dev),
                    "count": cur, "graph": torch.cuda.CUDAGraph(), "states": [self.state[p] for p in ps], "grads": list(grads),
                    "state_ptrs": [self.state[p]["exp_avg"].data_ptr() for p in ps]}
            params = [p.data for p in ps]
            ms, vs = [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps]
            torch.cuda.synchronize()
            with torch.cuda.graph(plan["graph"]):
                ops.adam_advance(plan["step_dev"], group["lr"], beta1, beta2, group["eps"], plan["hyper"])
                ops.adam_apply(params, list(grads), ms, vs, plan["hyper"], row_slot0=row_slot0)
            plan["exec"] = plan["graph"].raw_cuda_graph_exec()
            self._fast[key] = plan
        states = plan["states"]
        cur = states[0]["step"]
        for st, ptr in zip(states, plan["state_ptrs"]):
            if st["step"] != cur or st["exp_avg"].data_ptr() != ptr:  # load_state_dict / external steps changed the state under us
                self._fast.clear()
                return None
        if plan["count"] != cur:  # someone else advanced these parameters: resynchronise the device counter
            plan["step_dev"].fill_(cur)
            plan["count"] = cur
        return plan

    def fast_replay(self, plan, ps) -> None:
        from . import _lib
        from ._lib import check

        check(_lib.load().kgat_graph_launch(plan["exec"], torch.cuda.current_stream().cuda_stream), "adam graph launch")
        plan["count"] += 1
        for st in plan["states"]:
            st["step"] += 1
        torch.autograd.graph.increment_version(ps)  # the kernel wrote through raw pointers: tell autograd the data changed
        for p in ps:
            p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._step_impl()
        return loss

    def step_and_zero(self) -> None:
        """``step()`` followed by ``zero_grad()`` (set to None) -- what the reference's ``update_*_weights`` do
        (model.py:407-419) -- without torch.optim's per-call hook / profiler wrappers, which cost more host time than
        the 80 us of GPU work of a KG step.  Falls back to the wrapped methods as soon as any hook is registered."""
        from torch.optim import optimizer as _o

        if (getattr(self, "_optimizer_step_pre_hooks", None) or getattr(self, "_optimizer_step_post_hooks", None)
                or getattr(_o, "_global_optimizer_pre_hooks", None) or getattr(_o, "_global_optimizer_post_hooks", None)):
            self.step()
            self.zero_grad()
            return
        with torch.no_grad():
            self._step_impl()
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    def _step_impl(self) -> None:
        for group in self.param_groups:
            with_grad = [p for p in group["params"] if p.grad is not None]
            if all(not p.grad.is_sparse and p.grad.is_contiguous() for p in with_grad) and self._graphed_step(group, with_grad):
                continue
            by_step: dict[int, list] = {}
            for p in with_grad:
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
