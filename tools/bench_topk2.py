"""Times dmm_topk_edges_pruned against dmm_topk_edges on synthetic scores of the bench shapes (CUDA events, L2 flushed).
    python tools/bench_topk2.py [baby|sports|scaleout ...]
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth

DEV = 'cuda:0'
SHAPES = {"baby": (19445, 7050), "sports": (35598, 18357), "scaleout": (4096, 500000)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def chunk_max(sc):
    n_rows, n_cols = sc.shape
    nch = (n_cols + 31) // 32
    out = torch.empty((n_rows, nch), device=DEV)
    for r0 in range(0, n_rows, 2048):
        blk = sc[r0:r0 + 2048]
        pad = torch.full((blk.shape[0], nch * 32), float("-inf"), device=DEV)
        pad[:, :n_cols] = blk
        out[r0:r0 + 2048] = pad.view(blk.shape[0], nch, 32).amax(dim=2)
    return out


def run(name):
    U, I = SHAPES[name]
    inter = synth.interactions(U, I, seed=0)
    deg = np.diff(inter.indptr)
    ld = ops.pad_to(I, 4)
    buf = torch.randn((U, ld), device=DEV)
    buf.mul_(0.05)
    scores = buf[:, :I]
    cm = ops.cmax_buffer(U, I, DEV)
    cm[:, :(I + 31) // 32] = chunk_max(scores)
    dptr = torch.from_numpy(inter.indptr).to(DEV)
    order = ops.rows_long_first(dptr, 0, U, 32)
    items = torch.empty(int(inter.indptr[-1]), dtype=torch.int32, device=DEV)
    items2 = torch.empty_like(items)

    def timed(fn, label):
        ts = []
        for _ in range(6):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        print(f"{name:9s} {label:34s} {ms * 1e3:9.1f} us   {4.0 * U * I / ms / 1e6:9.1f} GB/s of 4*I*U", flush=True)
    timed(lambda: ops.topk_edges(scores, I, dptr, 0, None, items2, order=order), "whole-row (round 1)")
    timed(lambda: ops.topk_edges_pruned(scores, I, cm, dptr, 0, None, items, order=order), "pruned (chunk maxima)")
    assert torch.equal(items, items2)


for n in (sys.argv[1:] or ["baby", "sports", "scaleout"]):
    run(n)
