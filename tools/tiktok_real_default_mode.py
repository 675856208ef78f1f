"""Real-TikTok three-epoch runs (tests/golden/tiktok_real, conf/tiktok.toml) in the DEFAULT mode of the trainer -- device
generator, phases 1 and 3 replayed from CUDA graphs, bf16 -- next to the reference's 14-run ensemble: the parity tests gate
the seed-exact CPU-RNG / eager mode, this shows that the mode users actually run lands in the same distribution.
    python tools/tiktok_real_default_mode.py [repetitions]         DIFFMM_ADAM=dmm|foreach|torch_fused selects the Adam implementation"""
import glob
import json
import os
import sys
import tempfile

sys.path.insert(0, '.')
ROOT = os.path.abspath('.')
GOLD = os.path.join(ROOT, "tests", "golden", "tiktok_real")
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3      # repetitions of the same seed (they differ through the atomics)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle.gen_tiktok_golden import materialise  # noqa: E402
from diffmm_b200 import Main  # noqa: E402
from diffmm_b200.Conf import load_config  # noqa: E402

work = tempfile.mkdtemp(prefix="tiktok_def_")
materialise(GOLD, work)
os.chdir(work)
os.environ["DIFFMM_CPU_RNG"] = "0"
ref = []
for p in sorted(glob.glob(os.path.join(GOLD, "*.json"))):
    g = json.load(open(p))
    if isinstance(g, dict) and "epochs" in g and len(g["epochs"]) >= 3:
        ref.append([(e["test"]["Recall"], e["test"]["NDCG"], e["train"]["Loss"]) for e in g["epochs"][:3]])
ref = np.array(ref)
print(f"reference ensemble ({len(ref)} runs): " + " | ".join(
    f"R@20 {ref[:, e, 0].mean():.5f} +- {ref[:, e, 0].std(ddof=1):.5f} N@20 {ref[:, e, 1].mean():.5f} Loss {ref[:, e, 2].mean():.4f}" for e in range(3)))
runs = []
for s in range(n_seeds):
    cfg = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
    cfg.train.epoch = 3
    cfg.base.precision = "bf16"
    Main.seed_it(cfg.base.seed)
    h = Main.DataHandler(cfg)
    h.LoadData()
    coach = Main.Coach(h, cfg)
    coach.run()
    assert coach._use_graph()
    runs.append([(r["test"]["Recall"], r["test"]["NDCG"], r["train"]["Loss"]) for r in coach.history])
    print(f"seed {cfg.base.seed} adam={os.environ.get('DIFFMM_ADAM', 'dmm')}: " + " | ".join(
        f"R@20 {a:.5f} N@20 {b:.5f} Loss {c:.4f}" for a, b, c in runs[-1]), flush=True)
    del coach, h
    torch.cuda.empty_cache()
runs = np.array(runs)
for e in range(3):
    mu, sd = ref[:, e, 0].mean(), ref[:, e, 0].std(ddof=1)
    print(f"epoch {e}: ours mean R@20 {runs[:, e, 0].mean():.5f} ({100 * (runs[:, e, 0].mean() / mu - 1):+.2f} % of the reference mean, "
          f"reference sd {100 * sd / mu:.1f} %), Loss {runs[:, e, 2].mean():.4f} vs {ref[:, e, 2].mean():.4f}")
