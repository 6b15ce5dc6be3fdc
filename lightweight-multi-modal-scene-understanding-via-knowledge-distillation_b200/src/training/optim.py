"""AdamW over ONE flat fp32 buffer, stepped by a single CUDA kernel.

The reference steps ``torch.optim.AdamW`` over ~96 small parameter tensors
(src/training/trainer.py:56,90).  Here every parameter is a view into one flat
buffer (and every gradient a view into a second one), so that

  * the optimizer step is one launch of ``kdf_adamw_flat``;
  * data-parallel training all-reduces the gradients as ONE NCCL bucket
    (``flat_grad``), which at 2.1 MB is latency-bound, not bandwidth-bound;
  * ``zero_grad`` is one memset.

``state_dict()`` / ``load_state_dict()`` speak torch.optim.AdamW's per-parameter
format, so checkpoints written by the reference's Trainer load here and vice versa.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from .. import native as _native
from .. import ops


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        params = [p for p in params]
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      amsgrad=False, maximize=False, foreach=None, capturable=False,
                                      differentiable=False, fused=None, decoupled_weight_decay=True))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdamW supports a single parameter group")
        self._params: List[torch.nn.Parameter] = list(self.param_groups[0]["params"])
        dev = self._params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on CUDA parameters only (no CPU fallback)")
        for p in self._params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatAdamW needs float32 parameters on one CUDA device")
        self._offsets, total = [], 0
        for p in self._params:
            self._offsets.append(total)
            total += (p.numel() + 7) // 8 * 8                 # every fp32 view AND its bf16 shadow view stay 16-byte aligned
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._hyper = torch.zeros(2, dtype=torch.float32, device=dev)     # [lr, step]
        self._step = 0
        with torch.no_grad():
            for p, off in zip(self._params, self._offsets):
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + n].view(p.shape)
                p.grad = self.flat_grad[off:off + n].view(p.shape)
        # bf16 shadow of every parameter, kept current by the optimizer kernel: the tensor-core layers take their
        # operands as views of it (no per-step cast).  Writes that go through torch (load_state_dict's ``param.copy_``,
        # a broadcast into ``flat_param``) advance a version counter that ``shadow_of`` checks; the kernels' own
        # updates do not (and write the shadow themselves).
        self.flat_param_bf16 = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.refresh_shadow()
        for i, p in enumerate(self._params):
            p._kdf_shadow = (lambda i=i: self.shadow_of(i))

    def shadow_of(self, i: int) -> torch.Tensor:
        """bf16 view of parameter ``i``; the whole shadow is refreshed first when that parameter (or the flat buffer)
        was written through torch since the last refresh."""
        if self._params[i]._version != self._pver[i] or self.flat_param._version != self._shadow_version:
            self.refresh_shadow()
        off = self._offsets[i]
        return self.flat_param_bf16[off:off + self._params[i].numel()]

    @torch.no_grad()
    def refresh_shadow(self):
        self.flat_param_bf16.copy_(self.flat_param)
        self._shadow_version = self.flat_param._version
        self._pver = [p._version for p in self._params]

    # ------------------------------------------------------------------ stepping
    def zero_grad(self, set_to_none: bool = False):
        """Gradients are permanent views of ``flat_grad``; zeroing is one memset."""
        self.flat_grad.zero_()
        for p, off in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def detach_grads(self):
        """Drop the ``.grad`` views so that the next backward hands each gradient over as a fresh tensor instead of
        adding it into the view with one tiny kernel per parameter (96 launches for the weighted student);
        ``gather_grads_`` then moves all of them into ``flat_grad`` with one multi-tensor copy."""
        for p in self._params:
            p.grad = None

    @torch.no_grad()
    def gather_grads_(self, indices=None):
        """flat_grad <- the gradients autograd produced since ``detach_grads`` (zeros where a parameter got none);
        ``p.grad`` are views of ``flat_grad`` again afterwards.  ``indices`` restricts the move to those parameters
        (the Trainer gathers and all-reduces the early-arriving bucket while the rest of the backward still runs)."""
        srcs, dsts = [], []
        it = range(len(self._params)) if indices is None else indices
        for i in it:
            p, off = self._params[i], self._offsets[i]
            view = self.flat_grad[off:off + p.numel()].view(p.shape)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                srcs.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                dsts.append(view)
            p.grad = view
        if srcs:
            torch._foreach_copy_(dsts, srcs)

    def span(self, indices):
        """(lo, hi) element range of ``flat_grad`` that exactly covers the parameters ``indices`` when they are
        consecutive in the flat layout (padding included), else None."""
        idx = sorted(indices)
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            return None
        lo = self._offsets[idx[0]]
        hi = self._offsets[idx[-1] + 1] if idx[-1] + 1 < len(self._offsets) else self.flat_grad.numel()
        return lo, hi

    def set_hyper(self, lr: float, step: int):
        """Device-side (lr, step); called by step(), or by the caller before replaying a captured graph."""
        self._hyper.copy_(torch.tensor([lr, float(step)], dtype=torch.float32), non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, update_hyper: bool = True):
        if closure is not None:
            raise RuntimeError("FlatAdamW does not support closures")
        g = self.param_groups[0]
        if update_hyper:
            self._step += 1
            self.set_hyper(g["lr"], self._step)
        _native.bump_generation()                         # parameters change through the flat buffer (no _version bump)
        ops.adamw_flat_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._hyper,
                        g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], grad_scale, shadow=self.flat_param_bf16)

    # ------------------------------------------------------------------ torch.optim.AdamW-format checkpoints
    def state_dict(self):
        state = {}
        for i, (p, off) in enumerate(zip(self._params, self._offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self._step)),
                        "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self._params)))
        return {"state": state if self._step > 0 else {}, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("optimizer state does not match the parameter list")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        self.param_groups[0].setdefault("initial_lr", self.param_groups[0]["lr"])
        steps = set()
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(self._params, self._offsets)):
                st = sd["state"].get(i)
                if st is None:
                    continue
                n = p.numel()
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; cannot load into a flat optimizer")
        self._step = steps.pop() if steps else 0
