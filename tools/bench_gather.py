"""Times the first layer on the CSR rows (dmm_csr_gather_act / _split) at a benchmark shape (CUDA events, L2 flushed).
   python tools/bench_gather.py [baby|sports|tiktok]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth
DEV = 'cuda:0'
name = sys.argv[1] if len(sys.argv) > 1 else 'baby'
U, I, _ = synth.SHAPES[name]
H = 1024
heavy = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
inter = synth.interactions(U, I, seed=0, heavy_frac=heavy)
ptr, idx = torch.from_numpy(inter.indptr).to(DEV), torch.from_numpy(inter.indices).to(DEV)
g = torch.Generator(device=DEV).manual_seed(0)
wt = (torch.randn((I, H), device=DEV, generator=g) / 30).to(torch.bfloat16)
vals = torch.randn(idx.numel(), device=DEV, generator=g)
bias = torch.randn(H, device=DEV, generator=g)
h = torch.empty((U, H), dtype=torch.bfloat16, device=DEV)
z = torch.empty((U, H), device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
order = ops.rows_long_first(ptr, 0, U, 32)
plain = order.clone()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


nnz = idx.numel()
for label, o in (("per-slice warps", plain), ("rows divided by length", order)):
    for wl, v in (("weighted", vals), ("binary", None)):
        ms = timeit(lambda: ops.csr_gather_act(ptr, idx, U, I, wt, None, bias, 1, H, h, None, z_f32=z, order=o, vals=v))
        print(f"{name} heavy={heavy} {label:24s} {wl:9s}: {ms*1e3:7.1f} us  gathered {nnz*H*2/ms/1e6:6.0f} GB/s + written {U*H*6/ms/1e6:5.0f} GB/s")
