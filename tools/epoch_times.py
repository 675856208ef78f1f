"""Per-epoch phase seconds for a synthetic dataset shape, eager vs CUDA-graph joint training."""
import os, sys, tempfile, time
sys.path.insert(0, '.')
import torch
from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config
name = sys.argv[1] if len(sys.argv) > 1 else 'sports'
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_ep_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)
for graph in (False, True):
    cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = epochs; cfg.base.cuda_graph = graph
    Main.seed_it(0)
    h = Main.DataHandler(cfg); h.LoadData()
    coach = Main.Coach(h, cfg); coach.prepareModel()
    for ep in range(epochs):
        coach.phase_seconds = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        coach.trainEpoch(); coach.testEpoch(); torch.cuda.synchronize()
        print("graph" if graph else "eager", ep, round(time.perf_counter() - t0, 3), {k: round(v, 3) for k, v in coach.phase_seconds.items()},
              "mem GB", round(torch.cuda.memory_reserved() / 2**30, 2), flush=True)
    del coach, h
    torch.cuda.empty_cache()
