"""GPU: the one-launch Adam step (csrc/optim.cu, diffmm_b200.optim.FusedStepAdam) against torch.optim.Adam's capturable foreach
implementation -- the optimiser the graph-mode trainer used before -- on identical parameters and gradients: the update is
the same sequence of fp32 operations, so parameters and both moment buffers must agree BIT FOR BIT after every step."""
import pytest
import torch
from torch.optim.adam import Adam

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    shapes = [(1,), (7,), (64,), (1024, 1027), (300, 64), (3,), (129, 5)]
    return [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in shapes]


def test_fused_step_adam_is_bit_identical_to_torch_foreach_capturable():
    from diffmm_b200.optim import FusedStepAdam
    pa, pb = _params(0), _params(0)
    lr_a = torch.tensor(1e-3, device=DEV)
    lr_b = torch.tensor(1e-3, device=DEV)
    oa = FusedStepAdam(pa, lr=lr_a, weight_decay=0, capturable=True)
    ob = Adam(pb, lr=lr_b, weight_decay=0, capturable=True, foreach=True)
    g = torch.Generator(device=DEV).manual_seed(1)
    for it in range(40):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, device=DEV, generator=g) * (10.0 ** ((it % 7) - 4))
            if it % 5 == 0:
                gr[..., ::3] = 0.0                                  # exact zeros: rows of W1 of items nobody touched
            if it == 11:
                gr = gr * 1e-20                                      # denormal-range second moments
            x.grad, y.grad = gr.clone(), gr.clone()
        if it == 20:                                                 # the scheduler writes the tensor lr in place
            lr_a.fill_(3.7e-4)
            lr_b.fill_(3.7e-4)
        oa.step()
        ob.step()
        for k, (x, y) in enumerate(zip(pa, pb)):
            assert torch.equal(x, y), (it, k, float((x - y).abs().max()))
            sa, sb = oa.state[x], ob.state[y]
            assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), (it, k)
            assert torch.equal(sa["step"], sb["step"])
    # state_dict layout is torch's
    assert set(oa.state_dict()["state"][0].keys()) == set(ob.state_dict()["state"][0].keys())


def test_fused_step_adam_replays_from_a_cuda_graph():
    from diffmm_b200.optim import FusedStepAdam
    pa, pb = _params(2), _params(2)
    oa = FusedStepAdam(pa, lr=torch.tensor(2e-3, device=DEV), weight_decay=0, capturable=True)
    ob = Adam(pb, lr=torch.tensor(2e-3, device=DEV), weight_decay=0, capturable=True, foreach=True)
    g = torch.Generator(device=DEV).manual_seed(3)
    static = [torch.zeros_like(x) for x in pa]
    for x, s in zip(pa, static):
        x.grad = s
    s0 = torch.cuda.Stream()
    s0.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s0):
        for _ in range(3):                                           # warm-up steps outside the graph (state creation)
            for s, y in zip(static, pb):
                gr = torch.randn(s.shape, device=DEV, generator=g)
                s.copy_(gr)
                y.grad = gr.clone()
            oa.step()
            ob.step()
    torch.cuda.current_stream().wait_stream(s0)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        oa.step()
    for _ in range(5):
        for s, y in zip(static, pb):
            gr = torch.randn(s.shape, device=DEV, generator=g)
            s.copy_(gr)
            y.grad = gr.clone()
        graph.replay()
        ob.step()
    torch.cuda.synchronize()
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)


def test_fused_step_adam_eager_mode_is_bit_identical_to_torch_foreach():
    """The eager trainer's configuration: python-float lr (rewritten by the scheduler), step counters on the host."""
    from diffmm_b200.optim import FusedStepAdam
    pa, pb = _params(4), _params(4)
    oa = FusedStepAdam(pa, lr=1e-3, weight_decay=0)
    ob = Adam(pb, lr=1e-3, weight_decay=0, foreach=True)
    g = torch.Generator(device=DEV).manual_seed(5)
    for it in range(40):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, device=DEV, generator=g) * (10.0 ** ((it % 7) - 4))
            if it % 5 == 0:
                gr[..., ::3] = 0.0
            x.grad, y.grad = gr.clone(), gr.clone()
        if it == 20:
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 4.321e-4
        oa.step()
        ob.step()
        for k, (x, y) in enumerate(zip(pa, pb)):
            assert torch.equal(x, y), (it, k, float((x - y).abs().max()))
            sa, sb = oa.state[x], ob.state[y]
            assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), (it, k)
            assert float(sa["step"]) == float(sb["step"])
