"""torch.autograd.Function wrappers that give the C-ABI kernels forward + backward semantics.

PyTorch only provides the tape here; every contraction, SpMM and loss below runs in
libdiffmm_b200.so.  Backward formulas are stated next to each Function.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from . import ops

_PRECISIONS = ("bf16", "bf16x3")


def check_precision(p: str) -> str:
    if p not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {p!r}")
    return p


# ------------------------------------------------------------------------------------------------
# packed-weight cache: weights change once per optimiser step (tensor._version bumps), while the
# reverse-diffusion chain reuses them 5 x (#batches) times.
# ------------------------------------------------------------------------------------------------
_PACK_CACHE: dict = {}     # id(param) -> (weakref(param), version, {transpose: (hi, lo)})


def pack_any(src: torch.Tensor, transpose: bool, split: bool):
    """dmm_pack_bf16 of a 2-D fp32 tensor in either memory order: a column-major view (e.g. ``w.t()``)
    is packed by reading its row-major base with the opposite transpose flag."""
    if src.stride(1) == 1 and src.stride(0) >= src.shape[1]:
        return ops.pack_bf16(src, transpose=transpose, split=split)
    if src.stride(0) == 1 and src.stride(1) >= src.shape[0]:
        return ops.pack_bf16(src.t(), transpose=not transpose, split=split)
    return ops.pack_bf16(src.contiguous(), transpose=transpose, split=split)


def packed_weight(w: torch.Tensor, transpose: bool, split: bool):
    """bf16 hi/lo operand copy of ``w`` (or of w^T).  Only nn.Parameters are cached (stable storage,
    in-place updates bump ``_version``); the entry is validated by object identity, never by address."""
    if not isinstance(w, torch.nn.Parameter):
        return pack_any(w.detach(), transpose, split)
    ent = _PACK_CACHE.get(id(w))
    if ent is None or ent[0]() is not w or ent[1] != w._version:
        if len(_PACK_CACHE) > 256:
            for k in [k for k, v in _PACK_CACHE.items() if v[0]() is None]:
                del _PACK_CACHE[k]
        ent = (weakref.ref(w), w._version, {})
        _PACK_CACHE[id(w)] = ent
    packs = ent[2]
    have = packs.get(transpose)
    if have is None or (split and have[1] is None):
        # the lo part is only built when a caller asks for it (bf16 mode never does: a third less traffic)
        packs[transpose] = have = pack_any(w.detach(), transpose, split)
    hi, lo = have
    return hi, (lo if split else None)


def packed_weight_pair(w: torch.Tensor, split: bool):
    """Both orientations of a Parameter packed from ONE read (dmm_pack_bf16_pair) and entered into the cache, for
    weights that are needed both ways at the same version (Linear forward + input gradient; reverse chain).  Falls
    back to the separate packs for views that are not row-major."""
    if not isinstance(w, torch.nn.Parameter) or not (w.dim() == 2 and w.stride(1) == 1 and w.stride(0) >= w.shape[1]):
        return packed_weight(w, False, split), packed_weight(w, True, split)
    ent = _PACK_CACHE.get(id(w))
    if ent is None or ent[0]() is not w or ent[1] != w._version:
        ent = (weakref.ref(w), w._version, {})
        _PACK_CACHE[id(w)] = ent
    packs = ent[2]
    need = [t for t in (False, True) if packs.get(t) is None or (split and packs[t][1] is None)]
    if len(need) == 2:
        packs[False], packs[True] = ops.pack_bf16_pair(w.detach(), split=split)
    return packed_weight(w, False, split), packed_weight(w, True, split)


_CONST_CACHE: dict = {}    # (data_ptr, shape, stride) -> (weakref(tensor), version, {transpose: (hi, lo)})


def packed_const(x: torch.Tensor, transpose: bool, split: bool):
    """Packed copies of a constant input the CALLER vouches for (long-lived tensor, e.g. a registered feature
    matrix).  Keyed by storage address and validated by a weak reference to the very tensor object plus its
    version counter, so a recycled address or an in-place update can never return stale data."""
    key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()))
    ent = _CONST_CACHE.get(key)
    if ent is None or ent[0]() is not x or ent[1] != x._version:
        if len(_CONST_CACHE) > 64:
            for k in [k for k, v in _CONST_CACHE.items() if v[0]() is None]:
                del _CONST_CACHE[k]
        ent = (weakref.ref(x), x._version, {})
        _CONST_CACHE[key] = ent
    packs = ent[2]
    if transpose not in packs:
        packs[transpose] = pack_any(_rows(x.detach()), transpose, True)
    hi, lo = packs[transpose]
    return hi, (lo if split else None)


def _rows(t: torch.Tensor) -> torch.Tensor:
    """Row-major view/copy with unit inner stride (padded leading dimensions are fine)."""
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


def _new_out(M: int, N: int, device) -> torch.Tensor:
    ld = ops.pad_to(N, 4)
    buf = torch.empty((M, ld), dtype=torch.float32, device=device)
    return buf[:, :N] if ld != N else buf


class LinearTN(torch.autograd.Function):
    """y = act(x W^T + b) with x [M,K], W [N,K] (nn.Linear layout), on dmm_gemm_bf16_tn.

    backward (g = dL/dy, g' = g * (1 - y^2) for tanh):
        dx = g' W          -> gemm_tn(A = g' [M,N],   B = W^T [K,N])
        dW = g'^T x        -> gemm_tn(A = g'^T [N,M], B = x^T [K,M])
        db = sum_rows g'
    """

    @staticmethod
    def forward(ctx, x, weight, bias, act: int, precision: str, const_input: bool = False):
        split = precision == "bf16x3"
        x_obj = x                                  # the caller's tensor object: identity key of the constant cache
        x = _rows(x.detach())
        M, K = x.shape
        N = weight.shape[0]
        # const_input: x is a long-lived constant (the modality feature matrices, Model.py:47-58): its packed
        # copies are cached like the weights' instead of being rebuilt on every batch
        ctx.const_obj = x_obj if const_input else None
        x_hi, x_lo = packed_const(x_obj, False, split) if const_input else ops.pack_bf16(x, split=split)
        if ctx.needs_input_grad[0]:
            (w_hi, w_lo), _ = packed_weight_pair(weight, split)     # backward needs W^T at the same version: one read
        else:
            w_hi, w_lo = packed_weight(weight, False, split)
        y = _new_out(M, N, x.device)
        ops.gemm_bf16_tn(x_hi, x_lo, w_hi, w_lo, M, N, K, bias=bias.detach() if bias is not None else None, act=act,
                         out_f32=y)
        ctx.act, ctx.split, ctx.has_bias = act, split, bias is not None
        ctx.save_for_backward(x, weight, y if act else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        split = ctx.split
        g = _rows(g)
        if ctx.act == 1:
            g = g * (1.0 - y * y)
        M, K = x.shape
        N = weight.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            g_hi, g_lo = ops.pack_bf16(g, split=split)
            wt_hi, wt_lo = packed_weight(weight, True, split)
            dx = _new_out(M, K, g.device)
            ops.gemm_bf16_tn(g_hi, g_lo, wt_hi, wt_lo, M, K, N, out_f32=dx)
        if ctx.needs_input_grad[1]:
            gt_hi, gt_lo = ops.pack_bf16(g, transpose=True, split=split)
            xt_hi, xt_lo = packed_const(ctx.const_obj, True, split) if ctx.const_obj is not None else pack_any(x, True, split)
            dw = _new_out(N, K, g.device)
            ops.gemm_bf16_tn(gt_hi, gt_lo, xt_hi, xt_lo, N, K, M, out_f32=dw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = g.sum(0)
        return dx, dw, db, None, None, None


def linear_tn(x, weight, bias=None, act: int = 0, precision: str = "bf16", const_input: bool = False):
    return LinearTN.apply(x, weight, bias, act, check_precision(precision), const_input)


_SPMM_PRECISION = "bf16x3"     # propagation products: "bf16" = bf16 gather table (separable adjacencies), else fp32


def set_spmm_precision(precision: str) -> None:
    """Precision of the propagation products of this process (Coach sets it from ``base.precision``): "bf16" rounds the
    gathered operand to bf16 once per product (dmm_spmm_table_bf16 + dmm_spmm_norm_bf16: half the bytes through the L2,
    no value stream), "bf16x3" keeps the fp32 gather table (dmm_spmm_csr).  DIFFMM_SPMM_FP32=1 forces fp32."""
    global _SPMM_PRECISION
    _SPMM_PRECISION = check_precision(precision)


def _spmm_bf16(adj, x, precision=None) -> bool:
    """The bf16 gather table pays where the table does not sit in the L2 anyway: from DIFFMM_SPMM_BF16_MIN_NNZ stored
    entries (default 1 M: 131 vs 184 us at the ifashion shape; the shipped datasets, 0.16-0.6 M entries, are launch bound
    and keep the exact fp32 product, which costs one launch less)."""
    import os
    if (precision or _SPMM_PRECISION) != "bf16" or not adj.separable or x.shape[1] != 64 or not x.is_cuda:
        return False
    if os.environ.get("DIFFMM_SPMM_FP32", "0") == "1" or adj.n_nodes >= (1 << 25):
        return False
    return int(adj.nnz) >= int(os.environ.get("DIFFMM_SPMM_BF16_MIN_NNZ", "1000000"))


def _product(adj, x, out=None, row0=0, row1=None, precision=None):
    if _spmm_bf16(adj, x, precision):
        return ops.spmm_norm_bf16(adj, x, out=out, row0=row0, row1=row1)
    return ops.spmm(adj, x, out=out, row0=row0, row1=row1)


class SpMM(torch.autograd.Function):
    """y = A x for the symmetric normalised adjacency; dL/dx = A^T g = A g (same kernel)."""

    @staticmethod
    def forward(ctx, x, adj: ops.CsrAdj, precision=None):
        ctx.adj, ctx.precision = adj, precision
        return _product(adj, _rows(x.detach()), precision=precision)

    @staticmethod
    def backward(ctx, g):
        return _product(ctx.adj, _rows(g), precision=ctx.precision), None, None


class PartitionedSpMM(torch.autograd.Function):
    """y = A x with the rows of A (and of y) partitioned over the ranks of a process group (dist.PropPartition): every
    rank runs the CSR SpMM kernel on its own row blocks of the replicated adjacency and the blocks are all-gathered in
    place (NCCL over NVLink; one exchange per layer).  A is symmetric, so with a replicated upstream gradient g the
    backward dL/dx = A g is the same partitioned product: rows of A g on the owning rank + all-gather."""

    @staticmethod
    def _product(adj, x, part, precision=None):
        y = torch.empty((adj.n_nodes, x.shape[1]), dtype=torch.float32, device=x.device)
        if _spmm_bf16(adj, x, precision):
            table = ops.spmm_table_bf16(adj, x)       # one table for all of this rank's row blocks
            return part.product_(lambda r0, r1: ops.spmm_norm_bf16(adj, table=table, out=y, row0=r0, row1=r1), y)
        return part.product_(lambda r0, r1: ops.spmm(adj, x, out=y, row0=r0, row1=r1), y)

    @staticmethod
    def forward(ctx, x, adj: ops.CsrAdj, part, precision=None):
        ctx.adj, ctx.part, ctx.precision = adj, part, precision
        return PartitionedSpMM._product(adj, _rows(x.detach()), part, precision)

    @staticmethod
    def backward(ctx, g):
        return PartitionedSpMM._product(ctx.adj, _rows(g), ctx.part, ctx.precision), None, None, None


_PARTITION = None      # dist.PropPartition of the running trainer (None: single GPU)


def set_partition(part) -> None:
    """Row-partitions every propagation product of this process (Model.gcn_MM, the cross-layer CL layers of
    Coach._joint_step) over ``part.group``; None restores the single-GPU path."""
    global _PARTITION
    _PARTITION = part if (part is not None and part.world > 1) else None


class SpMMCat(torch.autograd.Function):
    """y = A [xa ; xb] without materialising the concatenation (bf16 propagation: the gather table is built straight from
    the two blocks, dmm_spmm_table_bf16); the gradient A g is handed back as two views."""

    @staticmethod
    def forward(ctx, xa, xb, adj: ops.CsrAdj):
        ctx.adj, ctx.na = adj, xa.shape[0]
        return ops.spmm_norm_bf16(adj, _rows(xa.detach()), x2=_rows(xb.detach()))

    @staticmethod
    def backward(ctx, g):
        gx = _product(ctx.adj, _rows(g), precision="bf16")
        return gx[:ctx.na], gx[ctx.na:], None


def spmm_cat(adj: ops.CsrAdj, xa: torch.Tensor, xb: torch.Tensor, precision=None) -> torch.Tensor:
    """spmm(adj, torch.cat([xa, xb])) (Model.py:89-93,110-114: user block, item block)."""
    if _PARTITION is None and xa.shape[1] == 64 and xb.shape[1] == 64 and _spmm_bf16(adj, xa, precision):
        return SpMMCat.apply(xa, xb, adj)
    return spmm(adj, torch.cat([xa, xb]), precision)


def spmm(adj: ops.CsrAdj, x: torch.Tensor, precision=None) -> torch.Tensor:
    """A x.  ``precision`` ("bf16" / "bf16x3"; None: the process default of set_spmm_precision) selects the bf16 gather
    table for separable adjacencies."""
    part = _PARTITION
    if part is not None and adj.n_nodes == part.n_nodes and (adj.n_users == part.n_users or adj.n_users == 0):
        return PartitionedSpMM.apply(x, adj, part, precision)
    return SpMM.apply(x, adj, precision)


class SpMMAxpy(torch.autograd.Function):
    """y = c (A x + x) in one launch (the SpMM epilogue adds the scaled operand): the tail of Model.gcn_MM,
    final = (m0 + A m0) + residual_weight (m0 + A m0) with c = 1 + residual_weight (Model.py:129-131).  A is symmetric:
    dL/dx = c (A g + g), the same call on the gradient."""

    @staticmethod
    def forward(ctx, x, adj: ops.CsrAdj, c: float):
        ctx.adj, ctx.c = adj, float(c)
        xd = _rows(x.detach())
        return ops.spmm(adj, xd, alpha=ctx.c, beta=ctx.c, z=xd)

    @staticmethod
    def backward(ctx, g):
        gd = _rows(g)
        return ops.spmm(ctx.adj, gd, alpha=ctx.c, beta=ctx.c, z=gd), None, None


class PartitionedSpMMAxpy(torch.autograd.Function):
    """SpMMAxpy with the rows partitioned over the ranks (dist.PropPartition): the same epilogue on every rank's row blocks
    (bit-identical to the single-GPU launch), then the in-place all-gather; backward = the same call on the gradient."""

    @staticmethod
    def _product(adj, x, part, c):
        y = torch.empty((adj.n_nodes, x.shape[1]), dtype=torch.float32, device=x.device)
        return part.product_(lambda r0, r1: ops.spmm(adj, x, alpha=c, beta=c, z=x, out=y, row0=r0, row1=r1), y)

    @staticmethod
    def forward(ctx, x, adj: ops.CsrAdj, part, c: float):
        ctx.adj, ctx.part, ctx.c = adj, part, float(c)
        return PartitionedSpMMAxpy._product(adj, _rows(x.detach()), part, ctx.c)

    @staticmethod
    def backward(ctx, g):
        return PartitionedSpMMAxpy._product(ctx.adj, _rows(g), ctx.part, ctx.c), None, None, None


def spmm_axpy(adj: ops.CsrAdj, x: torch.Tensor, c: float, precision=None) -> torch.Tensor:
    """c (A x + x); fused on the exact fp32 product (one GPU or row-partitioned), otherwise the plain composition."""
    if x.is_cuda and x.shape[1] == 64 and not _spmm_bf16(adj, x, precision):
        part = _PARTITION
        if part is None:
            return SpMMAxpy.apply(x, adj, c)
        if adj.n_nodes == part.n_nodes and (adj.n_users == part.n_users or adj.n_users == 0):
            return PartitionedSpMMAxpy.apply(x, adj, part, c)
    return c * (x + spmm(adj, x, precision))


class RowNormalize(torch.autograd.Function):
    """F.normalize(x) (dim 1, eps 1e-12) with a one-launch forward and a one-launch backward."""

    @staticmethod
    def forward(ctx, x):
        y, inv = ops.rownorm_fwd(_rows(x.detach()))
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        return ops.rownorm_bwd(y, inv, _rows(g))


def row_normalize(x: torch.Tensor) -> torch.Tensor:
    return RowNormalize.apply(x) if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 else \
        torch.nn.functional.normalize(x)


class ModalMix(torch.autograd.Function):
    """sum_m w[m] (y + lam z_m) (Model.py:116-119,125-127) with one launch forward and one backward (+ a fixed-order sum
    of the per-CTA partials for d/dw)."""

    @staticmethod
    def forward(ctx, w, lam: float, y, *zs):
        yd = y.detach().contiguous()
        zd = [z.detach().contiguous() for z in zs]
        wd = w.detach().contiguous()
        ctx.lam = float(lam)
        ctx.save_for_backward(wd, yd, *zd)
        return ops.modal_mix_fwd(yd, zd, wd, ctx.lam)

    @staticmethod
    def backward(ctx, g):
        wd, yd, *zd = ctx.saved_tensors
        gy, gzs, gw = ops.modal_mix_bwd(g.contiguous(), yd, zd, wd, ctx.lam)
        return (gw, None, gy, *gzs)


def modal_mix(w: torch.Tensor, lam: float, y: torch.Tensor, zs) -> torch.Tensor:
    return ModalMix.apply(w, lam, y, *zs)


class InfoNCEFn(torch.autograd.Function):
    """-mean_b log softmax_b(<n1_b, n2_.>/T)_b on gathered, L2-normalised rows (Utils/Utils.py:57-75).

    backward: with p = softmax rows, dL/dn1_i = sum_j (p_ij - d_ij) n2_j / (B T),
    dL/dn2_j = sum_i (p_ij - d_ij) n1_i / (B T), then the normalisation Jacobian
    (I - n n^T)/|x|, and a scatter-add over the (repeating) gather indices.
    """

    @staticmethod
    def forward(ctx, v1, v2, idx, temperature: float):
        v1d, v2d = _rows(v1.detach()), _rows(v2.detach())
        loss, saved = ops.infonce_fwd(v1d, v2d, idx, temperature)
        ctx.temperature = temperature
        ctx.save_for_backward(v1d, v2d, idx, *saved)
        return loss

    @staticmethod
    def backward(ctx, g):
        v1, v2, idx, lse, inv1, inv2 = ctx.saved_tensors
        g1, g2 = ops.infonce_bwd(v1, v2, idx, ctx.temperature, (lse, inv1, inv2), 1.0)
        d1 = d2 = None
        if ctx.needs_input_grad[0]:
            d1 = ops.scatter_add_rows(g1 * g, idx, torch.zeros_like(v1))
        if ctx.needs_input_grad[1]:
            d2 = ops.scatter_add_rows(g2 * g, idx, torch.zeros_like(v2))
        return d1, d2, None, None


class BPRFn(torch.autograd.Function):
    """mean_b -log(1e-5 + sigmoid(u_b.p_b - u_b.n_b)) on rows already gathered by the caller's
    indices (Utils/Utils.py:78-98); gradients come back per batch row."""

    @staticmethod
    def forward(ctx, u, p, n):
        B = u.shape[0]
        ar = torch.arange(B, device=u.device)
        ud, pd, nd = _rows(u.detach()), _rows(p.detach()), _rows(n.detach())
        # item table = [p; n]: pos rows 0..B-1, neg rows B..2B-1
        tab = torch.cat([pd, nd], 0)
        loss, grads = ops.bpr_fwd_bwd(ud, tab, ar, ar, ar + B, 1.0, True)
        ctx.save_for_backward(*grads)
        return loss

    @staticmethod
    def backward(ctx, g):
        gu, gp, gn = ctx.saved_tensors
        return gu * g, gp * g, gn * g


class JointLossFn(torch.autograd.Function):
    """BPR + every InfoNCE term of a joint-training step (Main.py:309,333,345-367) through dmm_bpr_infonce_fwd / _bwd:
    three launches forward, three backward, gradients scatter-added into one zeroed table per view.

    forward(spec, users, pos, neg, *tables) -> (bpr_loss, contrastive_total); spec = (problems, bpr_table, item_offset)
    with problems = [(i1, i2, 'u' | 'i', temperature, weight)] indexing `tables` ([N, 64] fp32 node tables)."""

    @staticmethod
    def forward(ctx, spec, users, pos, neg, *tables):
        problems, bpr_table, item_off = spec
        tabs = [_rows(t.detach()) for t in tables]
        B = users.numel()
        probs = [(i1, i2, users if kind == "u" else pos, 0 if kind == "u" else item_off, temp, w)
                 for (i1, i2, kind, temp, w) in problems]
        bpr = (bpr_table, item_off, users, pos, neg)
        losses, saved = ops.bpr_infonce_fwd(tabs, probs, bpr, B)
        ctx.probs, ctx.bpr, ctx.B, ctx.n_tables = probs, bpr, B, len(tables)
        ctx.save_for_backward(*tabs, *saved)
        P = len(probs)
        return losses[P], losses[P + 1]

    @staticmethod
    def backward(ctx, g_bpr, g_cl):
        n = ctx.n_tables
        tabs, saved = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        grads = [torch.zeros_like(t) if ctx.needs_input_grad[4 + k] else None for k, t in enumerate(tabs)]
        ops.bpr_infonce_bwd(list(tabs), ctx.probs, ctx.bpr, ctx.B, saved, g_cl.contiguous().float(), g_bpr.contiguous().float(),
                            grads)
        return (None, None, None, None, *grads)


def joint_losses(problems, bpr_table, item_offset, users, pos, neg, tables):
    return JointLossFn.apply((problems, bpr_table, item_offset), users, pos, neg, *tables)
