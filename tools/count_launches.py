"""Kernel launches per training batch (diffusion phase and joint phase), counted with torch.profiler (every CUDA kernel:
ours through the C ABI and ATen's), for the fused paths (default) and the per-op paths (DIFFMM_FUSED_TRAIN=0,
DIFFMM_FUSED_LOSS=0).  Eager mode (the CUDA-graph default replays the same kernels without host launches).
    python tools/count_launches.py [baby|tiktok]"""
import os
import sys
import tempfile
from collections import Counter

sys.path.insert(0, '.')
import torch
from torch.profiler import ProfilerActivity, profile

from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config

name = sys.argv[1] if len(sys.argv) > 1 else 'baby'
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_cl_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)


def count(fn):
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    c = Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
            c[ev.name.split("(")[0][:70]] += 1
    return c


for fused in ("1", "0"):
    os.environ["DIFFMM_FUSED_TRAIN"] = fused
    os.environ["DIFFMM_FUSED_LOSS"] = fused
    cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = 1; cfg.base.cuda_graph = False
    Main.seed_it(0)
    h = Main.DataHandler(cfg); h.LoadData()
    coach = Main.Coach(h, cfg); coach.prepareModel()
    coach.trainEpoch()                       # warm-up: allocator, adjacencies, caches
    rows = next(iter(h.diffusionLoader))[0]
    M = 3 if coach.has_audio else 2
    ts = [torch.randint(0, 5, (rows.shape[0],), device=coach.device) for _ in range(M)]
    acc = torch.zeros(3, dtype=torch.float64, device=coach.device)
    coach._diffusion_step(rows, ts, acc)
    cd = count(lambda: coach._diffusion_step(rows, ts, acc))
    users, pos, neg = next(iter(h.trainLoader))
    users, pos, neg = users.long().cuda(), pos.long().cuda(), neg.long().cuda()
    biadj = Main._as_csr(h.torchBiAdj)
    coach._joint_step(users, pos, neg, biadj)
    cj = count(lambda: coach._joint_step(users, pos, neg, biadj))
    label = "fused (default)" if fused == "1" else "per-op paths"
    print(f"{name} {label}: diffusion batch ({M} modalities) {sum(cd.values())} kernels; joint batch {sum(cj.values())} kernels")
    for title, c in (("diffusion", cd), ("joint", cj)):
        top = ", ".join(f"{k.split('::')[-1][:38]} x{v}" for k, v in c.most_common(12))
        print(f"   {title}: {top}")
    del coach, h
    torch.cuda.empty_cache()
