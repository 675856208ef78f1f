"""Times dmm_gemm_bf16_tn on the hidden-space step / score shapes of the rebuild (CUDA events, L2 flushed)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops
DEV = 'cuda:0'
M = int(sys.argv[1]) if len(sys.argv) > 1 else 19445
g = torch.Generator(device=DEV).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


H, I = 1024, 7050
h = (torch.randn((M, H), device=DEV, generator=g) * 0.5).to(torch.bfloat16)
h2 = torch.empty_like(h)
P = (torch.randn((H, H), device=DEV, generator=g) / 32).to(torch.bfloat16)
z = torch.randn((M, H), device=DEV, generator=g)
q = torch.randn(H, device=DEV, generator=g)
pb = torch.randn(H, device=DEV, generator=g)
ms = timeit(lambda: ops.gemm_bf16_tn(h, None, P, None, M, H, H, bias=q, alpha=0.9, beta=0.1, residual=z, out_f32=z, out_hi=h2,
                                     post_bias=pb, post_act=1))
print(f"hidden step {M}x{H}x{H}: {ms*1e3:.1f} us  {2.0*M*H*H/ms/1e9:.0f} TFLOP/s  {M*H*12/ms/1e6:.0f} GB/s epilogue+operand")
W2 = (torch.randn((I, H), device=DEV, generator=g) / 32).to(torch.bfloat16)
b2 = torch.randn(I, device=DEV, generator=g)
x = torch.empty((M, ops.pad_to(I, 32)), device=DEV)[:, :I]
ms = timeit(lambda: ops.gemm_bf16_tn(h, None, W2, None, M, I, H, bias=b2, alpha=1.0, out_f32=x))
print(f"scores      {M}x{I}x{H}: {ms*1e3:.1f} us  {2.0*M*I*H/ms/1e9:.0f} TFLOP/s")
W1 = (torch.randn((H, ops.pad_to(I, 64)), device=DEV, generator=g) / 80).to(torch.bfloat16)
W2t = (torch.randn((H, ops.pad_to(I, 64)), device=DEV, generator=g) / 80).to(torch.bfloat16)
Pf = torch.empty((H, H), device=DEV)
ms = timeit(lambda: ops.gemm_bf16_tn(W1, None, W2t, None, H, H, I, out_f32=Pf))
print(f"P = W1x W2  {H}x{H}x{I}: {ms*1e3:.1f} us  {2.0*H*H*I/ms/1e9:.0f} TFLOP/s")
Ph = torch.zeros((H, H), dtype=torch.bfloat16, device=DEV)
ms = timeit(lambda: ops.gemm_bf16_tn_splitk(W1[:, :I], W2t[:, :I], H, H, I, out_hi=Ph))
print(f"P split-K   {H}x{H}x{I}: {ms*1e3:.1f} us  {2.0*H*H*I/ms/1e9:.0f} TFLOP/s (contraction + slab reduce, bf16 operand out)")
ms = timeit(lambda: ops.gemm_bf16_tn(W1[:, :I], None, W2t[:, :I], None, H, H, I, out_hi=Ph))
print(f"P plain     {H}x{H}x{I}: {ms*1e3:.1f} us  (bf16 operand out)")
for I2 in (18357, 500000):
    A = (torch.randn((H, ops.pad_to(I2, 64)), device=DEV, generator=g) / 80).to(torch.bfloat16)
    B = (torch.randn((H, ops.pad_to(I2, 64)), device=DEV, generator=g) / 80).to(torch.bfloat16)
    ms = timeit(lambda: ops.gemm_bf16_tn_splitk(A[:, :I2], B[:, :I2], H, H, I2, out_hi=Ph), n=5)
    ms0 = timeit(lambda: ops.gemm_bf16_tn(A[:, :I2], None, B[:, :I2], None, H, H, I2, out_hi=Ph), n=5)
    print(f"P I={I2}: split-K {ms*1e3:.1f} us ({2.0*H*H*I2/ms/1e9:.0f} TFLOP/s)  plain {ms0*1e3:.1f} us ({2.0*H*H*I2/ms0/1e9:.0f} TFLOP/s)")
