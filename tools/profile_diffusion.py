"""CUDA-kernel time breakdown of phase 1 (diffusion training) batches via torch.profiler (baby shape)."""
import os, sys, tempfile
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config
name = 'baby'
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_prof_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)
cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = 3
Main.seed_it(0)
h = Main.DataHandler(cfg); h.LoadData()
coach = Main.Coach(h, cfg); coach.prepareModel()
coach.trainDiffusion(); torch.cuda.synchronize()
import time
t0 = time.perf_counter(); coach.trainDiffusion(); torch.cuda.synchronize(); print("phase 1 wall", time.perf_counter() - t0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    coach.trainDiffusion()
    torch.cuda.synchronize()
N = len(h.diffusionLoader)
ev = [e for e in prof.key_averages() if e.self_device_time_total > 0 and not e.key.startswith(("aten::", "Optimizer", "LinearTN", "autograd"))]
tot = sum(e.self_device_time_total for e in ev)
print(f"GPU kernel time per batch: {tot / N / 1e3:.3f} ms over {sum(e.count for e in ev) / N:.0f} kernels/batch ({N} batches)")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:30]:
    print(f"{e.self_device_time_total / N:9.1f} us/batch  x{e.count / N:5.1f}  {e.key[:120]}")
