"""Times dmm_topk_edges on baby-shape scores for several degree distributions (CUDA events, L2 flushed)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth
DEV = 'cuda:0'
U, I = 19445, 7050
inter = synth.interactions(U, I, seed=0)
deg = np.diff(inter.indptr)
buf = torch.randn((U, 7072), device=DEV); buf.mul_(0.05); scores = buf[:, :I]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def run(k, label, ordered=False):
    ptr = np.zeros(U + 1, dtype=np.int64); np.cumsum(k, out=ptr[1:])
    dptr = torch.from_numpy(ptr).to(DEV)
    order = ops.rows_long_first(dptr, 0, U, 32) if ordered else None
    items = torch.empty(int(ptr[-1]), dtype=torch.int32, device=DEV)
    ts = []
    for it in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.topk_edges(scores, I, dptr, 0, None, items, order=order); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    print(f"{label:40s} {ms*1e3:8.1f} us  {4.0*U*I/ms/1e6:8.1f} GB/s   k: mean {k.mean():.1f} max {k.max()} >256: {(k>256).sum()} >128: {(k>128).sum()}", flush=True)
run(deg, "bench degrees (1% heavy tail)")
run(deg, "bench degrees, rows with k > 32 first", ordered=True)
run(np.minimum(deg, 256), "clipped to 256")
run(np.minimum(deg, 100), "clipped to 100")
run(np.minimum(deg, 8), "clipped to 8")
run(np.full(U, 1), "k = 1")
run(np.full(U, 20), "k = 20")
run(np.full(U, 64), "k = 64")
