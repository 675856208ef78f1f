"""Times dmm_spmm_csr on the ifashion-shaped graph (CUDA events)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth
DEV = 'cuda:0'
name = sys.argv[1] if len(sys.argv) > 1 else 'ifashion'
U, I, _ = synth.SHAPES[name]
inter = synth.interactions(U, I, seed=0)
ptr = torch.from_numpy(inter.indptr).to(DEV); idx = torch.from_numpy(inter.indices).to(DEV)
adj = ops.build_norm_adj(ptr, idx, U, I)
N = U + I
x = torch.randn((N, 64), device=DEV); y = torch.empty_like(x)
deg = (adj.ptr[1:] - adj.ptr[:-1]).cpu().numpy()
print("N", N, "nnz", adj.nnz, "max row", deg.max(), "rows>256:", (deg > 256).sum(), "nnz in rows>256:", deg[deg > 256].sum())
for _ in range(3): ops.spmm(adj, x, out=y)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.spmm(adj, x, out=y); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts)); by = 8.0 * adj.nnz + 8.0 * (N + 1) + 2.0 * N * 64 * 4
print(f"spmm {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic (one call per event pair)")
# back to back: 20 calls inside one event pair (launch overhead of the host hidden behind the queue)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.spmm(adj, x, out=y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"spmm {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic (20 calls back to back)")
