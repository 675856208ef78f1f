"""SpMM on the ifashion-shaped graph and the shipped shapes (CUDA events around one Python call, median of 10; the small
shapes are dominated by the ~15 us call overhead): fp32 gather table (dmm_spmm_csr), bf16 gather table with the separable
normalisation (dmm_spmm_norm_bf16) with and without the table pass inside the timed call."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from diffmm_b200 import ops, synth
DEV = 'cuda:0'


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run(U, I, label):
    inter = synth.interactions(U, I, seed=0)
    ptr = torch.from_numpy(inter.indptr).to(DEV); idx = torch.from_numpy(inter.indices).to(DEV)
    adj = ops.build_norm_adj(ptr, idx, U, I)
    N = U + I
    x = torch.randn((N, 64), device=DEV); y = torch.empty_like(x)
    by = 8.0 * adj.nnz + 8.0 * (N + 1) + 2.0 * N * 64 * 4
    ms = timeit(lambda: ops.spmm(adj, x, out=y))
    want = y.clone()
    print(f"{label}: N {N} nnz {adj.nnz}  fp32 table {ms*1e3:7.1f} us = {by/ms/1e6:6.0f} GB/s algorithmic")
    table = ops.spmm_table_bf16(adj, x)
    ms4 = timeit(lambda: ops.spmm_norm_bf16(adj, table=table, out=y))
    err = float((y - want).abs().max() / want.abs().max())
    print(f"{label}:   v3 separable bf16 (table given) {ms4*1e3:7.1f} us = {by/ms4/1e6:6.0f} GB/s algorithmic   max err / scale {err:.2e}")
    ms5 = timeit(lambda: ops.spmm_norm_bf16(adj, x, out=y))
    print(f"{label}:   v3 separable bf16 (+ table pass) {ms5*1e3:7.1f} us = {by/ms5/1e6:6.0f} GB/s algorithmic")
    ms6 = timeit(lambda: ops.spmm_table_bf16(adj, x, table=table))
    print(f"{label}:   table pass alone {ms6*1e3:7.1f} us")


run(300000, 80000, "ifashion")
run(100000, 27000, "third   ")
run(35598, 18357, "sports  ")
run(19445, 7050, "baby    ")
run(9308, 6710, "tiktok  ")
