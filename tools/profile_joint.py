"""CUDA-kernel time breakdown of phase 3 (joint training) batches via torch.profiler (baby shape)."""
import os, sys, tempfile
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config
from diffmm_b200.Model import _as_csr
name = 'baby'
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_prof_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)
cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = 3
Main.seed_it(0)
h = Main.DataHandler(cfg); h.LoadData()
coach = Main.Coach(h, cfg); coach.prepareModel()
coach.trainEpoch(); torch.cuda.synchronize()
biadj = _as_csr(h.torchBiAdj)
it = iter(h.trainLoader)
def step():
    u, p, n = next(it)
    coach._joint_step(u.long().cuda(), p.long().cuda(), n.long().cuda(), biadj)
for _ in range(3): step()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"GPU time per step: {tot / N / 1e3:.3f} ms over {sum(e.count for e in ev) / N:.0f} kernels/step")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total / N:9.1f} us/step  x{e.count / N:5.1f}  {e.key[:110]}")

kern = [e for e in ev if not e.key.startswith(("aten::", "Optimizer", "InfoNCEFn", "SpMM", "LinearTN", "_SignNoise", "autograd", "Memcpy", "Memset", "BPR", "bpr", "ScatterRows")) or e.key.startswith(("Memcpy", "Memset"))]
print("--- by launch count (kernels only) ---")
print(f"kernels/step: {sum(e.count for e in kern) / N:.0f}")
for e in sorted(kern, key=lambda e: -e.count)[:45]:
    print(f"x{e.count / N:6.1f}  {e.self_device_time_total / N:8.1f} us/step  {e.key[:120]}")
