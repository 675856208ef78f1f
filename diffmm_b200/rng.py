"""Random draws of the hot path, kept as the SAME torch calls, shapes and order as the reference
(SURVEY.md Appendix B).  By default noise comes from the device generator exactly like the reference
on a GPU (Model.py:400, Main.py:320).  With DIFFMM_CPU_RNG=1 (parity tests) the draws are made on the
CPU generator and copied over, which reproduces the reference's CPU run bit for bit."""
import os

import torch


def cpu_rng() -> bool:
    return os.environ.get("DIFFMM_CPU_RNG", "0") == "1"


def randn_like(x: torch.Tensor) -> torch.Tensor:
    if cpu_rng():
        return torch.randn(tuple(x.shape), dtype=x.dtype).to(x.device)
    return torch.randn(tuple(x.shape), dtype=x.dtype, device=x.device)


def rand_like(x: torch.Tensor) -> torch.Tensor:
    if cpu_rng():
        return torch.rand(tuple(x.shape), dtype=x.dtype).to(x.device)
    return torch.rand(tuple(x.shape), dtype=x.dtype, device=x.device)
