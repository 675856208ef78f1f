import sys, numpy as np, torch
sys.path.insert(0, '.')
from diffmm_b200 import ops
from oracle import diffmm_oracle as O
DEV = 'cuda:0'
def run(n_cols, ld, n_rows=8, kmax=40, k_special=True):
    rng = np.random.default_rng(n_cols)
    scores = (rng.standard_normal((n_rows, n_cols)) * 0.05).astype(np.float32)
    k = np.minimum(rng.integers(1, kmax, n_rows), n_cols)
    if k_special:
        k[1], k[2] = n_cols, min(n_cols, 603)
    buf = torch.full((n_rows, ld), float('nan'), device=DEV)
    buf[:, :n_cols] = torch.from_numpy(scores).to(DEV)
    ptr = np.zeros(n_rows + 1, dtype=np.int64); np.cumsum(k, out=ptr[1:])
    E = int(ptr[-1])
    items = torch.full((E,), -1, dtype=torch.int32, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    try:
        ops.topk_edges(buf[:, :n_cols], n_cols, torch.from_numpy(ptr).to(DEV), 100, None, items, status)
        torch.cuda.synchronize()
        want = np.concatenate(O.topk_edges(scores, k))
        ok = np.array_equal(items.cpu().numpy(), want)
        print(n_cols, ld, 'special' if k_special else 'plain', 'OK' if ok else 'MISMATCH', flush=True)
    except Exception as e:
        print(n_cols, ld, 'special' if k_special else 'plain', 'ERROR', str(e)[:80], flush=True)
        sys.exit(1)
for a in sys.argv[1:]:
    n, ld, sp = a.split(':')
    run(int(n), int(ld), k_special=bool(int(sp)))
