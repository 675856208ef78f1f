"""2+ GPU NCCL check (run under torchrun): the user-sharded rebuild + edge all-gather reproduces the single-GPU
edge lists bit for bit, and the row-partitioned propagation (local SpMM on a row block + all-gather of the X
blocks, dist.py) reproduces the full SpMM.  Prints 'MULTIGPU OK' on rank 0."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as td
from diffmm_b200 import dist as ddist, ops, rebuild, synth
from diffmm_b200.Conf import Config
from diffmm_b200.Model import Denoise, GaussianDiffusion

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
td.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
U, I, H = 3001, 1500, 256
inter = synth.interactions(U, I, seed=1)
cfg = Config(); cfg.base.precision = "bf16"; cfg.base.denoise_dim = f"[{H}]"; cfg.data.user_num, cfg.data.item_num = U, I
torch.manual_seed(0)
gd = GaussianDiffusion(cfg).to(dev)
dens = {m: Denoise([I, H], [H, I], cfg).to(dev) for m in ("image", "text")}
ptr = torch.from_numpy(inter.indptr).to(dev); idx = torch.from_numpy(inter.indices).to(dev)
single = rebuild.rebuild_edges(gd, dens, ptr, idx, U, I, 0, "bf16")
sharded = rebuild.rebuild_modal_adj(gd, dens, ptr, idx, U, I, 0, "bf16", group=td.group.WORLD)
full = {m: ops.build_norm_adj(ptr, v, U, I) for m, v in single.items()}
ok = True
for m in dens:
    ok &= bool(torch.equal(full[m].idx, sharded[m].idx)) and bool(torch.equal(full[m].val, sharded[m].val))
# row-partitioned propagation, 3 layers
adj = full["image"]
N = U + I
torch.manual_seed(1)
x = torch.randn((N, 64), device=dev)
want = x
for _ in range(3):
    want = ops.spmm(adj, want)
blocks = ddist.row_blocks(N, world)
a, b = blocks[rank]
cur = x
for _ in range(3):
    y = torch.zeros((N, 64), device=dev)
    ops.spmm(adj, cur, out=y, row0=a, row1=b)
    cur = ddist.allgather_rows(y[a:b].contiguous(), blocks)
ok &= bool(torch.allclose(cur, want, rtol=1e-6, atol=1e-7))
flag = torch.tensor([1 if ok else 0], device=dev)
td.all_reduce(flag, op=td.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU OK" if int(flag.item()) == 1 else "MULTIGPU MISMATCH", "world", world, flush=True)
td.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
