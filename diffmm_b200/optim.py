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
    """torch.optim.Adam whose step is ONE launch (dmm_adam_step / dmm_adam_step_host).  Needs CUDA fp32 dense parameters,
    weight_decay 0, amsgrad False, maximize False and either capturable=True with a tensor learning rate on the device (the
    graph-mode trainer) or capturable=False with a python-float learning rate (the eager trainer): the two operation
    sequences of torch's foreach implementation, reproduced bit for bit.  Anything else falls back to the parent's step."""

    def _supported(self, group) -> bool:
        base = (group["weight_decay"] == 0 and not group["amsgrad"] and not group["maximize"] and not group["differentiable"]
                and not group.get("fused") and 0.5 < group["betas"][0] < 1.0
                and not isinstance(group["betas"][0], torch.Tensor) and not isinstance(group["betas"][1], torch.Tensor))
        if not base:
            return False
        if group["capturable"]:       # graph mode: device-resident step counters and learning rate
            return isinstance(group["lr"], torch.Tensor) and group["lr"].is_cuda and group["lr"].dtype == torch.float32
        return not isinstance(group["lr"], torch.Tensor)      # eager mode: python-float lr, step counters on the host

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
            if ok and not group["capturable"]:
                ok = all(st.is_cpu for st in steps) and len({float(st.item()) for st in steps}) == 1
            if not ok:                                          # nothing has been touched yet: the parent takes the whole step
                return super().step(closure)
            work.append((group, params, grads, exp_avgs, exp_avg_sqs, steps))
        for group, params, grads, exp_avgs, exp_avg_sqs, steps in work:
            grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
            n = len(params)
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])     # noqa: E731
            numel = (C.c_int64 * n)(*[p.numel() for p in params])
            beta1, beta2 = group["betas"]
            if group["capturable"]:
                torch._foreach_add_(steps, 1)                   # every parameter keeps its own (equal) step counter
                _lib.call("dmm_adam_step", ops._ctx(params[0]), n, arr(params), arr(grads), arr(exp_avgs), arr(exp_avg_sqs), numel,
                          ops._p(steps[0]), ops._p(group["lr"]), float(beta1), float(beta2), float(group["eps"]), ops._stream())
            else:
                # torch's non-capturable foreach sequence: the bias corrections are python doubles of the host-side step count
                for st in steps:
                    st += 1
                t = float(steps[0].item())                      # a CPU tensor: no device sync
                step_size = (group["lr"] / (1 - beta1 ** t)) * -1
                bc2_sqrt = (1 - beta2 ** t) ** 0.5
                _lib.call("dmm_adam_step_host", ops._ctx(params[0]), n, arr(params), arr(grads), arr(exp_avgs), arr(exp_avg_sqs),
                          numel, float(step_size), float(bc2_sqrt), float(beta1), float(beta2), float(group["eps"]), ops._stream())
            # the kernel wrote the parameters behind autograd's back: bump their version counters like torch's in-place
            # foreach ops do (the operand-pack caches of the trainer are keyed by them)
            torch.autograd.graph.increment_version(params)
        return None
