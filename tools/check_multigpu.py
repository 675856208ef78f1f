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
# partitioned propagation in the PRODUCT path: Model.gcn_MM + the cross-layer CL layers through autograd.spmm, values
# and parameter gradients against the single-GPU path (replicated parameters, row-partitioned products)
from diffmm_b200 import autograd as ag
from diffmm_b200.Model import Model
cfg.data.image_feat_dim, cfg.data.text_feat_dim = 48, 32
torch.manual_seed(2)
feats = torch.randn((I, 48), device=dev), torch.randn((I, 32), device=dev)
model = Model(cfg, feats[0], feats[1]).to(dev)
biadj = ops.build_norm_adj(ptr, idx, U, I)
wu, wi = torch.randn((U, 64), device=dev), torch.randn((I, 64), device=dev)


def prop_loss():
    out = model.gcn_MM(biadj, full["image"], full["text"])
    e = torch.cat([model.u_embs, model.i_embs])
    cl = 0.0
    for _ in range(3):
        e = ag.spmm(biadj, e)
        cl = cl + e
    return (out.u_final_embs * wu).sum() + (out.i_final_embs * wi).sum() + (out.u_text_embs * wu).sum() * 0.1 + (cl[:U] * wu).sum() * 0.01


def grads():
    model.zero_grad(set_to_none=True)
    l = prop_loss()
    l.backward()
    return l.detach(), [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]


ag.set_partition(None)
l1, g1 = grads()
ag.set_partition(ddist.PropPartition(U, I, td.group.WORLD))
l2, g2 = grads()
ag.set_partition(None)
ok_prop = bool(torch.allclose(l1, l2, rtol=1e-5)) and len(g1) == len(g2) and all(
    bool(torch.allclose(a_, b_, rtol=1e-4, atol=1e-6 * float(a_.abs().max() + 1e-30))) for a_, b_ in zip(g1, g2))
if rank == 0:
    print("partitioned gcn_MM + CL layers: loss", float(l1), float(l2), "grads equal:", ok_prop, flush=True)
ok &= ok_prop
flag = torch.tensor([1 if ok else 0], device=dev)
td.all_reduce(flag, op=td.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU OK" if int(flag.item()) == 1 else "MULTIGPU MISMATCH", "world", world, flush=True)
td.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
