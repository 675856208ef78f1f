"""Times dmm_infonce_fwd / dmm_infonce_bwd at the joint-training batch shape (B = 1024, D = 64), CUDA events."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops
DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device=DEV).manual_seed(0)
v1 = torch.randn((26495, 64), device=DEV, generator=g)
v2 = v1 + 0.5 * torch.randn((26495, 64), device=DEV, generator=g)
idx = torch.randint(0, 26495, (B,), device=DEV, generator=g)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


loss, saved = ops.infonce_fwd(v1, v2, idx, 0.5)
print(f"B={B} fwd {timeit(lambda: ops.infonce_fwd(v1, v2, idx, 0.5)):.1f} us   "
      f"bwd {timeit(lambda: ops.infonce_bwd(v1, v2, idx, 0.5, saved)):.1f} us   (fwd 2*B*B*64 = {2*B*B*64/1e6:.0f} MFLOP, bwd 3x)")
