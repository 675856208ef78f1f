"""GPU tests at BASELINE.json's full sizes through size-independent properties (the CPU oracle cannot reach
them in seconds): top-k against torch.topk values and tie rules, the rebuilt adjacency against its defining
identities, the reverse chain against a torch fp32 restatement on the same device, and a row wider than the
register / shared-memory paths of the top-k (scale-out config: >= 10^5 items)."""
import numpy as np
import pytest
import torch

from oracle import diffmm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from diffmm_b200 import ops as o
    return o


def _baby():
    from diffmm_b200 import synth
    U, I, _ = synth.SHAPES["baby"]
    inter = synth.interactions(U, I, seed=11)
    return U, I, inter


def test_topk_full_size_matches_torch_topk_values(ops):
    U, I, inter = _baby()
    g = torch.Generator(device=DEV).manual_seed(5)
    buf = torch.randn((U, 7072), device=DEV, generator=g)
    scores = buf[:, :I]
    ptr = torch.from_numpy(inter.indptr).to(DEV)
    deg = np.diff(inter.indptr)
    items = torch.full((int(inter.indptr[-1]),), -1, dtype=torch.int32, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.topk_edges(scores, I, ptr, 0, None, items, status)
    assert int(status.item()) == 0
    got = items.cpu().numpy()
    kmax = int(deg.max())
    vals, _ = torch.topk(scores, kmax, dim=1)                          # continuous scores: no ties
    vals = vals.cpu().numpy()
    sc = scores.cpu().numpy()
    for u in np.random.default_rng(0).choice(U, 2000, replace=False):
        k = int(deg[u])
        mine = got[inter.indptr[u]:inter.indptr[u + 1]]
        assert len(mine) == k and (np.diff(mine) > 0).all()            # ascending, unique columns
        np.testing.assert_array_equal(np.sort(sc[u, mine])[::-1], vals[u, :k])
    # every slot written, every column in range
    assert (got >= 0).all() and (got < I).all()


def test_adjacency_full_size_properties(ops):
    U, I, inter = _baby()
    ptr = torch.from_numpy(inter.indptr).to(DEV)
    idx = torch.from_numpy(inter.indices).to(DEV)
    adj = ops.build_norm_adj(ptr, idx, U, I)
    N, E = U + I, int(idx.numel())
    ap, ai, av = adj.ptr.cpu().numpy(), adj.idx.cpu().numpy(), adj.val.cpu().numpy()
    assert ap[0] == 0 and ap[-1] == 2 * E + N and (np.diff(ap) >= 1).all()
    rows = np.repeat(np.arange(N), np.diff(ap))
    # strictly ascending columns inside every row, exactly one self loop per row
    same = rows[1:] == rows[:-1]
    assert (ai[1:][same] > ai[:-1][same]).all()
    assert (ai == rows).sum() == N
    # symmetric pattern and values: sort the transposed triplets and compare
    o1 = np.lexsort((ai, rows))
    o2 = np.lexsort((rows, ai))
    np.testing.assert_array_equal(rows[o1], ai[o2])
    np.testing.assert_array_equal(ai[o1], rows[o2])
    np.testing.assert_array_equal(av[o1], av[o2])
    # val = d_r^-1/2 d_c^-1/2 with d = row length (self loop included), rounded like the reference
    d = np.diff(ap).astype(np.float64)
    want = ((d[rows] ** -0.5) * 1.0) * (d[ai] ** -0.5)
    np.testing.assert_array_equal(av, want.astype(np.float32))
    # SpMM with the all-ones vector = row sums (planned long rows included)
    x = torch.ones((N, 64), device=DEV)
    y = ops.spmm(adj, x).cpu().numpy()
    rs = np.bincount(rows, weights=av.astype(np.float64), minlength=N)
    np.testing.assert_allclose(y[:, 0], rs, rtol=3e-6)
    np.testing.assert_allclose(y[:, 63], rs, rtol=3e-6)


def test_chain_full_size_against_torch_fp32():
    """bf16x3 chain on the whole baby matrix vs the same arithmetic in torch fp32 on the device."""
    from diffmm_b200.Conf import Config
    from diffmm_b200.Model import Denoise, GaussianDiffusion
    from diffmm_b200.rebuild import denoise_chain
    U, I, inter = _baby()
    cfg = Config()
    cfg.base.precision = "bf16x3"
    cfg.data.user_num, cfg.data.item_num = U, I
    torch.manual_seed(3)
    gd = GaussianDiffusion(cfg).to(DEV)
    den = Denoise([I, 1024], [1024, I], cfg).to(DEV)
    ptr = torch.from_numpy(inter.indptr).to(DEV)
    idx = torch.from_numpy(inter.indices).to(DEV)
    rows = torch.arange(4096, 4096 + 512, device=DEV)
    with torch.no_grad():
        got = denoise_chain(gd, den, csr=(ptr, idx), row_ids=rows, n_rows=512).clone()
        x0 = torch.zeros((512, I), device=DEV)
        for r, u in enumerate(rows.tolist()):
            x0[r, idx[ptr[u]:ptr[u + 1]].long()] = 1.0
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        x = x0
        W1, b1 = den.in_layers[0].weight, den.in_layers[0].bias
        W2, b2 = den.out_layers[0].weight, den.out_layers[0].bias
        for i in range(gd.steps - 1, -1, -1):
            t = torch.full((512,), i, device=DEV)
            half = den.time_emb_dim // 2
            freqs = torch.exp(-np.log(10000.0) * torch.arange(half, device=DEV, dtype=torch.float32) / half)
            ang = t[:, None].float() * freqs[None]
            temb = den.emb_layer(torch.cat([torch.cos(ang), torch.sin(ang)], -1))
            h = torch.tanh(torch.cat([x, temb], -1) @ W1.t() + b1)
            pred = h @ W2.t() + b2
            x = float(np.float32(gd._h_coef1[i])) * pred + float(np.float32(gd._h_coef2[i])) * x
        torch.backends.cuda.matmul.allow_tf32 = prev
    err = (got - x).abs().max().item()
    assert err < 5e-5 * max(1.0, x.abs().max().item()), err


def test_chain_and_topk_wide_rows(ops):
    """Scale-out width: 120000 items (top-k generic kernels, GEMM with 469 column blocks) on a few users,
    checked against the numpy oracle."""
    from diffmm_b200.Conf import Config
    from diffmm_b200.Model import Denoise, GaussianDiffusion
    from diffmm_b200.rebuild import rebuild_edges
    U, I, H = 48, 120000, 256
    rng = np.random.default_rng(2)
    deg = rng.integers(1, 40, U)
    deg[0] = 700
    cols = [np.sort(rng.choice(I, int(k), replace=False)) for k in deg]
    indptr = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = np.concatenate(cols).astype(np.int32)
    cfg = Config()
    cfg.base.precision = "bf16x3"
    cfg.base.denoise_dim = f"[{H}]"
    cfg.data.user_num, cfg.data.item_num = U, I
    torch.manual_seed(4)
    gd = GaussianDiffusion(cfg).to(DEV)
    den = Denoise([I, H], [H, I], cfg).to(DEV)
    items = rebuild_edges(gd, {"m": den}, torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV), U, I,
                          precision="bf16x3")["m"].cpu().numpy()
    f = lambda t: t.detach().cpu().numpy()  # noqa: E731
    params = dict(emb_w=f(den.emb_layer.weight), emb_b=f(den.emb_layer.bias), w1=f(den.in_layers[0].weight),
                  b1=f(den.in_layers[0].bias), w2=f(den.out_layers[0].weight), b2=f(den.out_layers[0].bias),
                  gate_w=f(den.gate_layer.weight), gate_b=f(den.gate_layer.bias))
    sched = O.make_schedule(cfg.hyper.noise_scale, cfg.hyper.noise_min, cfg.hyper.noise_max, cfg.hyper.steps)
    x0 = np.zeros((U, I), dtype=np.float32)
    for u in range(U):
        x0[u, cols[u]] = 1.0
    view = O.generate_view(sched, params, x0, 0)
    want = O.topk_edges(view, deg)
    hit = sum(len(set(items[indptr[u]:indptr[u + 1]].tolist()) & set(w.tolist())) for u, w in enumerate(want))
    assert hit / indptr[-1] >= 0.995, hit / indptr[-1]
    for u in range(U):
        mine = items[indptr[u]:indptr[u + 1]]
        assert len(mine) == deg[u] and (np.diff(mine) > 0).all()
