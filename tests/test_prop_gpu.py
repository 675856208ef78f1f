"""GPU: the fused element-wise glue of Model.gcn_MM (csrc/prop.cu) against the per-op torch expressions of the reference
(Model.py:89-93 F.normalize, :116-127 the modality mix, :129-131 the residual tail): values and gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ag():
    from diffmm_b200 import autograd
    return autograd


def test_row_normalize_matches_f_normalize(ag):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((1000, 64), device=DEV, generator=g)
    x[3] = 0.0                     # clamped at eps: y = 0, gradient g / eps
    x[7] *= 1e-20
    up = torch.randn((1000, 64), device=DEV, generator=g)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya, yb = ag.row_normalize(a), F.normalize(b)
    # the same division; the norm itself may differ in its last bit (torch reduces the squares in another order)
    np.testing.assert_allclose(ya.detach().cpu().numpy(), yb.detach().cpu().numpy(), rtol=3e-7, atol=0)
    (ya * up).sum().backward()
    (yb * up).sum().backward()
    ok = torch.ones(1000, dtype=torch.bool, device=DEV)
    ok[[3, 7]] = False
    np.testing.assert_allclose(a.grad[ok].cpu().numpy(), b.grad[ok].cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(a.grad[3].cpu().numpy(), b.grad[3].cpu().numpy(), rtol=1e-6)      # both: g / eps
    for D in (32, 100):
        x2 = torch.randn((37, D), device=DEV, generator=g)
        np.testing.assert_allclose(ag.row_normalize(x2).cpu().numpy(), F.normalize(x2).cpu().numpy(), rtol=3e-7, atol=0)


@pytest.mark.parametrize("M", [2, 3])
def test_modal_mix_matches_the_per_op_expression(ag, M):
    g = torch.Generator(device=DEV).manual_seed(M)
    N, lam = 2701, 0.6
    y0 = torch.randn((N, 64), device=DEV, generator=g)
    z0 = [torch.randn((N, 64), device=DEV, generator=g) for _ in range(M)]
    w0 = torch.randn(M, device=DEV, generator=g)
    up = torch.randn((N, 64), device=DEV, generator=g)

    def leafs():
        return (y0.clone().requires_grad_(True), [z.clone().requires_grad_(True) for z in z0], w0.clone().requires_grad_(True))

    y, zs, w = leafs()
    weight = torch.softmax(w, -1)
    out = ag.modal_mix(weight, lam, y, zs)
    (out * up).sum().backward()
    y2, zs2, w2 = leafs()
    weight2 = torch.softmax(w2, -1)
    ref = None
    for m in range(M):
        aware = y2 + lam * zs2[m]
        ref = weight2[m] * aware if ref is None else ref + weight2[m] * aware
    (ref * up).sum().backward()
    assert torch.equal(out, ref)                                  # same operations in the same order
    np.testing.assert_allclose(y.grad.cpu().numpy(), y2.grad.cpu().numpy(), rtol=1e-5, atol=1e-6)
    for a, b in zip(zs, zs2):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(w.grad.cpu().numpy(), w2.grad.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(w2.grad.abs().max()))


def test_spmm_axpy_matches_the_composition(ag):
    from diffmm_b200 import ops, synth
    inter = synth.interactions(900, 700, seed=3)
    ptr = torch.from_numpy(inter.indptr).to(DEV)
    idx = torch.from_numpy(inter.indices).to(DEV)
    adj = ops.build_norm_adj(ptr, idx, 900, 700)
    g = torch.Generator(device=DEV).manual_seed(5)
    x0 = torch.randn((1600, 64), device=DEV, generator=g)
    up = torch.randn((1600, 64), device=DEV, generator=g)
    a = x0.clone().requires_grad_(True)
    b = x0.clone().requires_grad_(True)
    c = 1.2
    ya = ag.spmm_axpy(adj, a, c, "bf16x3")
    t = b + ag.spmm(adj, b, "bf16x3")
    yb = t + (c - 1.0) * t
    (ya * up).sum().backward()
    (yb * up).sum().backward()
    scale = float(yb.abs().max())
    np.testing.assert_allclose(ya.detach().cpu().numpy(), yb.detach().cpu().numpy(), rtol=2e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=2e-5, atol=2e-6 * float(b.grad.abs().max()))
