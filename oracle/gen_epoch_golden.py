"""TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference trainer (Main.Coach.run) for two epochs on a
tiny synthetic tiktok-named dataset on CPU and records the per-epoch loss dicts and Recall/NDCG.
The dataset files and the result are committed under tests/golden/epoch_run/ so that the GPU test can
replay exactly the same run through diffmm_b200 (DIFFMM_CPU_RNG=1 reproduces the CPU RNG stream).

    python oracle/gen_epoch_golden.py
"""
import json
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from diffmm_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "epoch_run")
U, I = 300, 200
FEATS = dict(image=16, text=24, audio=8)
OVER = {"base.denoise_dim": "[32]", "base.seed": 1818, "train.batch": 128, "train.test_batch": 128,
        "train.epoch": 2, "hyper.noise_scale": 0.5, "hyper.noise_degree": 1.5, "hyper.cross_cl_rate": 0.5,
        "hyper.sim_weight": 0.01, "train.reg": 1e-4}


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    inter = synth.interactions(U, I, seed=7, mean_deg=5.0, heavy_frac=0.02)
    synth.write_dataset(OUT, "tiktok", inter, synth.features(I, FEATS, seed=7))
    ref = ref_shim.load_reference()
    cfg = ref_shim.make_config(ref, "tiktok", **OVER)
    ref.Main.config = cfg
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        ref.Main.seed_it(cfg.base.seed)
        handler = ref.DataHandler.DataHandler(cfg)
        handler.LoadData()
        coach = ref.Main.Coach(handler, cfg)
        results = []
        orig_train, orig_test = coach.trainEpoch, coach.testEpoch

        def train():
            r = orig_train()
            results.append({"train": {k: float(v) for k, v in r.items()}})
            return r

        def test():
            r = orig_test()
            results[-1]["test"] = {k: float(v) for k, v in r.items()}
            return r

        coach.trainEpoch, coach.testEpoch = train, test
        coach.run()
    finally:
        os.chdir(cwd)
    shutil.rmtree(os.path.join(OUT, "logs"), ignore_errors=True)
    with open(os.path.join(OUT, "result.json"), "w") as f:
        json.dump({"overrides": OVER, "users": U, "items": I, "epochs": results,
                   "torch": torch.__version__, "numpy": np.__version__}, f, indent=1)
    print(json.dumps(results, indent=1))
    os.system(f"du -sh {OUT}")


if __name__ == "__main__":
    main()
