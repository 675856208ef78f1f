"""Statistical check of the CUDA-graph joint-training mode: per-epoch losses and Recall@20 for eager runs with two
device-RNG seeds and for the graph mode (same data, same init, same negatives)."""
import os, sys, tempfile
sys.path.insert(0, '.')
import torch
from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config
name = 'baby'
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_cmp_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)
def run(tag, graph, cuda_seed, epochs=4):
    cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = epochs; cfg.base.cuda_graph = graph
    Main.seed_it(0)
    h = Main.DataHandler(cfg); h.LoadData()
    coach = Main.Coach(h, cfg); coach.prepareModel()
    torch.cuda.manual_seed_all(cuda_seed)
    out = []
    for ep in range(epochs):
        r = coach.trainEpoch(); t = coach.testEpoch()
        out.append((round(r["Loss"], 4), round(r["BPR Loss"], 4), round(r["CL loss"], 4), round(float(t["Recall"]), 4)))
    print(tag, out, flush=True)
run("eager seed 0 ", False, 0)
run("eager seed 1 ", False, 1)
run("eager seed 2 ", False, 2)
run("graph seed 0 ", True, 0)
run("graph seed 1 ", True, 1)
