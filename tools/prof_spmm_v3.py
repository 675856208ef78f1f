"""One product of each SpMM path on the ifashion-shaped graph, for `ncu --set full` (profiles/r02_prof_spmm_v3.txt)."""
import sys
import torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth
DEV = 'cuda:0'
U, I = 300000, 80000
inter = synth.interactions(U, I, seed=0)
adj = ops.build_norm_adj(torch.from_numpy(inter.indptr).to(DEV), torch.from_numpy(inter.indices).to(DEV), U, I)
x = torch.randn((U + I, 64), device=DEV)
y = torch.empty_like(x)
for _ in range(2):
    ops.spmm_norm_bf16(adj, x, out=y)
    ops.spmm(adj, x, out=y)
torch.cuda.synchronize()
