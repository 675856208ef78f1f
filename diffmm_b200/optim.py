"""Adam with a one-launch update (csrc/optim.cu) behind torch.optim.Adam's interface (Main.py:92-110: the reference's
optimisers).  State layout, hyper-parameters, state_dict and the scheduler interplay are torch's: only
``step()`` differs, and only in HOW the same fp32 operations are scheduled -- one pass over p / g / m / v instead of the
fourteen of the capturable foreach implementation, bit-identical results (tests/test_optim_gpu.py)."""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim.adam import Adam

from . import _lib, ops


class FusedStepAdam(Adam):
    """torch.optim.Adam(capturable=True) whose step is dmm_adam_step.  Needs CUDA fp32 dense parameters, weight_decay 0,
    amsgrad False, maximize False and a tensor learning rate on the device (the configuration of the graph-mode trainer);
    anything else falls back to the parent's step."""

    def _supported(self, group) -> bool:
        return (group["weight_decay"] == 0 and not group["amsgrad"] and not group["maximize"] and group["capturable"]
                and not group["differentiable"] and isinstance(group["lr"], torch.Tensor) and group["lr"].is_cuda
                and group["lr"].dtype == torch.float32 and 0.5 < group["betas"][0] < 1.0)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None or not all(self._supported(g) for g in self.param_groups):
            return super().step(closure)
        check = getattr(self, "_accelerator_graph_capture_health_check", None) or getattr(self, "_cuda_graph_capture_health_check", None)
        if check is not None:
            check()
        work = []
        for group in self.param_groups:
            params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps = [], [], [], [], [], []
            self._init_group(group, params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps)
            if not params:
                continue
            ok = all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.dtype == torch.float32 and not g.is_sparse
                     for p, g in zip(params, grads))
            if not ok:                                          # nothing has been touched yet: the parent takes the whole step
                return super().step(closure)
            work.append((group, params, grads, exp_avgs, exp_avg_sqs, steps))
        for group, params, grads, exp_avgs, exp_avg_sqs, steps in work:
            grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
            torch._foreach_add_(steps, 1)                       # every parameter keeps its own (equal) step counter
            n = len(params)
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])     # noqa: E731
            numel = (C.c_int64 * n)(*[p.numel() for p in params])
            beta1, beta2 = group["betas"]
            _lib.call("dmm_adam_step", ops._ctx(params[0]), n, arr(params), arr(grads), arr(exp_avgs), arr(exp_avg_sqs), numel,
                      ops._p(steps[0]), ops._p(group["lr"]), float(beta1), float(beta2), float(group["eps"]), ops._stream())
        return None
