"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the DiffMM hot path.

This is the parity oracle for diffmm_b200's CUDA kernels.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import it; the product package never does (it fails loudly when the
CUDA library is missing).

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the unmodified reference executed in the
build container: ``oracle/gen_golden.py`` writes ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function below against them.

Every function cites the reference file:line (paths relative to the reference
root) whose arithmetic it restates.  fp32 unless noted; the schedule and the
SNR tail are fp64 exactly like the reference.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- schedule (a3)
def get_betas(noise_scale, noise_min, noise_max, steps):
    """Model.py:239-250."""
    start = noise_scale * noise_min
    end = noise_scale * noise_max
    variance = np.linspace(start, end, steps, dtype=np.float64)
    alpha_bar = 1 - variance
    betas = [1 - alpha_bar[0]]
    for i in range(1, steps):
        betas.append(min(1 - alpha_bar[i] / alpha_bar[i - 1], 0.999))
    return np.array(betas, dtype=np.float64)


def make_schedule(noise_scale, noise_min, noise_max, steps, beta_fixed=True):
    """Model.py:232-237,252-275.  All fp64.  The reference concatenates an fp32
    ``torch.tensor([1.0])`` with fp64 cumprod (type-promoted to fp64)."""
    betas = get_betas(noise_scale, noise_min, noise_max, steps)
    if beta_fixed:
        betas[0] = 0.0001
    alphas = 1.0 - betas
    ac = np.cumprod(alphas)
    ac_prev = np.concatenate([[1.0], ac[:-1]])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=betas,
        alphas_cumprod=ac,
        alphas_cumprod_prev=ac_prev,
        sqrt_alphas_cumprod=np.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
        posterior_variance=post_var,
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def snr(sched, t):
    """Model.py:380-383 (fp64)."""
    ac = sched["alphas_cumprod"]
    return ac[t] / (1 - ac[t] + 1e-8)


def snr_weight(sched, t):
    """Model.py:410-412: SNR(max(t-1,0)) - SNR(t), forced to 1 where t == 0 (fp64)."""
    t = np.asarray(t)
    tm1 = np.clip(t - 1, 0, None)
    w = snr(sched, tm1) - snr(sched, t)
    return np.where(t == 0, 1.0, w)


# --------------------------------------------------------------------------- Denoise (a2)
def time_embedding(timesteps, d_emb):
    """Model.py:196-201: [cos(t f), sin(t f)], f_j = exp(-ln(1e4) j / (d//2)); zero-pad if d odd."""
    half = d_emb // 2
    freqs = np.exp(-F32(math.log(10000)) * np.arange(half, dtype=F32) / F32(half)).astype(F32)
    temp = np.asarray(timesteps).astype(F32)[:, None] * freqs[None, :]
    emb = np.concatenate([np.cos(temp), np.sin(temp)], axis=-1).astype(F32)
    if d_emb % 2:
        emb = np.concatenate([emb, np.zeros_like(emb[:, :1])], axis=-1)
    return emb


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def denoise_forward(p, x_t, timesteps, modal_feat=None):
    """Model.py:183-220 for single in/out layers (in_dims=[I,H], out_dims=[H,I]).

    ``p`` holds fp32 arrays: emb_w (d,d), emb_b, w1 (H, I+d), b1, w2 (I, H), b2,
    gate_w (64,64), gate_b.  Multi-layer lists are handled when w1/w2 are lists.
    """
    x_t = np.asarray(x_t, dtype=F32)
    d = p["emb_w"].shape[0]
    temb = time_embedding(timesteps, d) @ p["emb_w"].T + p["emb_b"]
    if modal_feat is not None:
        proj = x_t @ modal_feat                                   # Model.py:205
        gate = _sigmoid(proj @ p["gate_w"].T + p["gate_b"])       # :206
        x_t = x_t + (proj * gate) @ modal_feat.T                  # :207-208
    h = np.concatenate([x_t, temb.astype(F32)], axis=-1)          # :210
    w1s = p["w1"] if isinstance(p["w1"], (list, tuple)) else [p["w1"]]
    b1s = p["b1"] if isinstance(p["b1"], (list, tuple)) else [p["b1"]]
    w2s = p["w2"] if isinstance(p["w2"], (list, tuple)) else [p["w2"]]
    b2s = p["b2"] if isinstance(p["b2"], (list, tuple)) else [p["b2"]]
    for w, b in zip(w1s, b1s):
        h = np.tanh(h @ w.T + b)                                  # :211-213
    for i, (w, b) in enumerate(zip(w2s, b2s)):
        h = h @ w.T + b                                           # :214-215
        if i != len(w2s) - 1:
            h = np.tanh(h)                                        # :216-218
    return h.astype(F32)


# --------------------------------------------------------------------------- q_sample (a4)
def l2_normalize_rows(x, eps=1e-12):
    """torch.nn.functional.normalize(p=2, dim=1, eps=1e-12)."""
    n = np.sqrt((x.astype(F32) ** 2).sum(axis=1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


def forward_cal_xt(sched, x0, timesteps, noise=None, randn=None):
    """Model.py:324-341.  ``noise`` explicit, or the default
    ``sign(x0) * normalize(randn)`` with the gaussian draw passed as ``randn``."""
    x0 = np.asarray(x0, dtype=F32)
    if noise is None:
        noise = np.sign(x0) * l2_normalize_rows(np.asarray(randn, dtype=F32))
    a = sched["sqrt_alphas_cumprod"][timesteps].astype(F32)[:, None]           # :339,352
    b = sched["sqrt_one_minus_alphas_cumprod"][timesteps].astype(F32)[:, None]
    return (a * x0 + b * noise.astype(F32)).astype(F32)


# --------------------------------------------------------------------------- p_sample (a7)
def p_mean(sched, p, x_t, timesteps):
    """Model.py:357-378 (mean only; the variance terms are computed but unused)."""
    pred = denoise_forward(p, x_t, timesteps)                                   # :365, no gate
    c1 = sched["posterior_mean_coef1"][timesteps].astype(F32)[:, None]
    c2 = sched["posterior_mean_coef2"][timesteps].astype(F32)[:, None]
    return (c1 * pred + c2 * x_t).astype(F32)


def generate_view(sched, p, x_start, sampling_step, randn=None):
    """Model.py:300-322.  Deterministic reverse loop i = S-1 .. 0."""
    x_start = np.asarray(x_start, dtype=F32)
    steps = len(sched["betas"])
    B = x_start.shape[0]
    if sampling_step == 0:
        x_t = x_start
    else:
        x_t = forward_cal_xt(sched, x_start, np.full(B, sampling_step - 1), randn=randn)
    for i in range(steps - 1, -1, -1):
        x_t = p_mean(sched, p, x_t, np.full(B, i))
    return x_t


# --------------------------------------------------------------------------- losses of a5
def cosine_similarity_rows(a, b, eps=1e-8):
    """F.cosine_similarity(dim=-1, eps=1e-8): a.b / (max(|a|,eps) * max(|b|,eps))."""
    na = np.maximum(np.sqrt((a * a).sum(-1, dtype=F32)), F32(eps))
    nb = np.maximum(np.sqrt((b * b).sum(-1, dtype=F32)), F32(eps))
    return ((a * b).sum(-1, dtype=F32) / (na * nb)).astype(F32)


def training_losses(sched, p, x_start, i_embs, modal_feat, timesteps, noise, reg, sim_weight):
    """Model.py:385-428 with the two RNG draws (timesteps :397, noise :400) injected.
    Returns the (B,) fp64 per-row loss."""
    x_start = np.asarray(x_start, dtype=F32)
    x_t = forward_cal_xt(sched, x_start, timesteps, noise=noise)               # :401
    out = denoise_forward(p, x_t, timesteps, modal_feat)                       # :404
    mse = ((out - x_start) ** 2).mean(axis=-1, dtype=F32)                      # :407-408
    w = snr_weight(sched, timesteps)                                           # :410-412 (fp64)
    rec = w * mse.astype(np.float64)                                           # :413
    um = out @ modal_feat                                                      # :416
    ui = x_start @ i_embs                                                      # :417
    sim = F32(1) - cosine_similarity_rows(um, ui)                              # :418
    reg_loss = l2_reg_loss(reg, [i_embs])                                      # :421
    return rec + sim.astype(np.float64) * sim_weight + np.float64(reg_loss) * reg   # :425


# --------------------------------------------------------------------------- Utils losses (a12,a13)
def l2_reg_loss(reg, embeddings):
    """Utils/Utils.py:45-54 (fp32 accumulation)."""
    s = F32(0)
    for e in embeddings:
        s = F32(s + (np.asarray(e, dtype=F32) ** 2).sum(dtype=F32))
    return F32(s * F32(reg))


def log_softmax_rows(s):
    m = s.max(axis=1, keepdims=True)
    z = s - m
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True, dtype=F32))


def info_nce(v1, v2, idx, temperature, b_cos=True):
    """Utils/Utils.py:57-75."""
    a = np.asarray(v1, dtype=F32)[idx]
    b = np.asarray(v2, dtype=F32)[idx]
    if a.shape != b.shape:
        raise ValueError("InfoNCE expected the same shape for two views")
    if b_cos:
        a, b = l2_normalize_rows(a), l2_normalize_rows(b)
    s = (a @ b.T) / F32(temperature)
    return F32(-np.diag(log_softmax_rows(s)).mean(dtype=F32))


def bpr_loss(u, p, n):
    """Utils/Utils.py:78-98: mean(-log(10e-6 + sigmoid(u.p - u.n)))."""
    pos = (u * p).sum(1, dtype=F32)
    neg = (u * n).sum(1, dtype=F32)
    return F32((-np.log(F32(10e-6) + _sigmoid(pos - neg))).mean(dtype=F32))


# --------------------------------------------------------------------------- top-k rebuild (a8)
def topk_edges(scores, k_per_row):
    """Main.py:224-230.  For row r emit the ``k_per_row[r]`` largest entries.

    Stated tie-break (torch.topk leaves it unspecified): value descending, then
    item index ascending.  NaN never occurs on this path.  Returns a list of
    int64 index arrays, each sorted ASCENDING by item index (the adjacency does
    not depend on intra-row order; the CUDA kernel emits the same order).
    """
    out = []
    scores = np.asarray(scores, dtype=F32)
    for r in range(scores.shape[0]):
        k = int(k_per_row[r])
        if k <= 0:
            out.append(np.zeros(0, dtype=np.int64))
            continue
        order = np.lexsort((np.arange(scores.shape[1]), -scores[r].astype(np.float64)))
        out.append(np.sort(order[:k]).astype(np.int64))
    return out


def user_degrees(indptr):
    """DataHandler.py:133-143 on a binary CSR train matrix."""
    return np.diff(np.asarray(indptr)).astype(np.int64)


# --------------------------------------------------------------------------- adjacency (a9)
def normalized_adj_csr(u_idx, i_idx, n_users, n_items):
    """Coach.makeTorchAdj (Main.py:113-116) -> DataHandler.makeTorchAdj (DataHandler.py:69-93)
    -> normalizeAdj (:53-66), restated without scipy.

    A = [[0,R],[R^T,0]] binarised, + I, then D^-1/2 A D^-1/2 with D = row sums
    (self loop included), values ``(d_r^-1/2 * a) * d_c^-1/2`` in fp64 cast to
    fp32.  Returned canonical (row-major sorted, duplicates merged) CSR:
    indptr int64 (N+1), indices int64, vals fp32.
    """
    N = n_users + n_items
    u = np.asarray(u_idx, dtype=np.int64)
    i = np.asarray(i_idx, dtype=np.int64) + n_users
    rows = np.concatenate([u, i, np.arange(N)])
    cols = np.concatenate([i, u, np.arange(N)])
    key = np.unique(rows * N + cols)            # binarise: duplicates collapse
    rows, cols = key // N, key % N
    deg = np.bincount(rows, minlength=N).astype(np.float64)
    dinv = np.where(deg > 0, deg ** (-0.5), 0.0)
    vals = ((dinv[rows] * 1.0) * dinv[cols]).astype(F32)
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=N), out=indptr[1:])
    return indptr, cols.astype(np.int64), vals


def spmm_csr(indptr, indices, vals, x):
    """torch.sparse.mm(A, X) (Model.py:90 etc.) as a sequential CSR row reduction in fp32."""
    x = np.asarray(x, dtype=F32)
    N = len(indptr) - 1
    y = np.zeros((N, x.shape[1]), dtype=F32)
    # vectorised segmented sum (np.add.reduceat needs non-empty segments; self loops guarantee that)
    prod = vals[:, None].astype(F32) * x[indices]
    nz = np.diff(indptr) > 0
    y[nz] = np.add.reduceat(prod, indptr[:-1][nz], axis=0)
    return y


# --------------------------------------------------------------------------- gcn_MM (a10)
def softmax1d(w):
    w = np.asarray(w, dtype=F32)
    e = np.exp(w - w.max())
    return (e / e.sum(dtype=F32)).astype(F32)


def gcn_mm(u_embs, i_embs, feats, lin_w, lin_b, modal_weight, adj, modal_adjs,
           modal_adj_weight, residual_weight):
    """Model.py:60-134.  ``adj`` / ``modal_adjs[m]`` are (indptr, indices, vals) CSR triples;
    ``feats[m]`` raw modality features, ``lin_w[m]``/``lin_b[m]`` the Linear(feat_dim, 64).

    Reproduces the aliasing at Model.py:129-131 (``final_embs = modal_embs`` then two
    in-place adds): final = (m0 + A m0) + rw * (m0 + A m0).
    Returns dict(final, modal=[Z_m...]) with N x 64 arrays (users first).
    """
    w = softmax1d(modal_weight)                                           # :87
    base = np.concatenate([u_embs, i_embs]).astype(F32)
    y = spmm_csr(*adj, base)                                              # :110-114,122-123
    zs = []
    m0 = None
    for m, f in enumerate(feats):
        fm = (f.astype(F32) @ lin_w[m].T + lin_b[m]).astype(F32)          # :84-85,103
        z = spmm_csr(*modal_adjs[m], np.concatenate([u_embs, l2_normalize_rows(fm)]))  # :89-93,104-105
        zs.append(z)
        aware = y + F32(modal_adj_weight) * z                             # :116-117,125
        m0 = w[m] * aware if m0 is None else m0 + w[m] * aware            # :119,127
    t = m0 + spmm_csr(*adj, m0)                                           # :130 (in place on m0)
    final = t + F32(residual_weight) * t                                  # :131 (m0 aliases final)
    return dict(final=final.astype(F32), modal=zs)


# --------------------------------------------------------------------------- cross-layer CL (a11)
def cl_propagate(adj, u_embs, i_embs, rand_uniform, noise_degree, layers=3):
    """Main.py:315-330.  ``rand_uniform`` is the list of the three torch.rand_like draws.
    Returns (mean of the perturbed layers, layer-1 output)."""
    e = np.concatenate([u_embs, i_embs]).astype(F32)
    outs = []
    for k in range(layers):
        e = spmm_csr(*adj, e)                                             # :319
        e = e + np.sign(e) * l2_normalize_rows(rand_uniform[k]) * F32(noise_degree)  # :320-321
        outs.append(e)
    mean = (np.stack(outs).mean(axis=0, dtype=F32)).astype(F32)           # :325
    return mean, outs[0]


# --------------------------------------------------------------------------- eval (f1, adjacent)
def eval_topk(user_embs, item_embs, users, train_indptr, train_indices, topk):
    """Main.py:410-411: scores = U_b I^T * (1-mask) - mask*1e8, top-`topk` indices per user
    (value desc, index asc)."""
    out = np.zeros((len(users), topk), dtype=np.int64)
    for r, u in enumerate(users):
        s = (item_embs @ user_embs[u]).astype(F32)
        seen = train_indices[train_indptr[u]:train_indptr[u + 1]]
        s[seen] = s[seen] * 0 - F32(1e8)
        order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))
        out[r] = order[:topk]
    return out


def recall_ndcg(top_idxs, test_items, topk):
    """Main.py:422-448 for a list of users; returns summed (recall, ndcg, precision)."""
    R = N = P = 0.0
    for rec, its in zip(top_idxs, test_items):
        rec = list(rec)
        tst = len(its)
        max_dcg = sum(1.0 / math.log2(loc + 2) for loc in range(min(tst, topk)))
        hits = dcg = 0.0
        for it in its:
            if it in rec:
                hits += 1
                dcg += 1.0 / math.log2(rec.index(it) + 2)
        R += hits / tst
        N += dcg / max_dcg
        P += hits / topk
    return R, N, P
