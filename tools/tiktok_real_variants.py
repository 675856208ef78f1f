"""Real-TikTok epoch runs (tests/golden/tiktok_real) through Coach under the arithmetic variants of this repo, next to
the reference's own runs: how far the trajectory moves when only the arithmetic changes.
    python tools/tiktok_real_variants.py [epochs]"""
import itertools
import json
import os
import sys
import tempfile

sys.path.insert(0, '.')
ROOT = os.path.abspath('.')
GOLD = os.path.join(ROOT, "tests", "golden", "tiktok_real")
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
import torch  # noqa: E402
from oracle.gen_tiktok_golden import materialise  # noqa: E402
from diffmm_b200 import Main  # noqa: E402
from diffmm_b200.Conf import load_config  # noqa: E402

work = tempfile.mkdtemp(prefix="tiktok_var_")
materialise(GOLD, work)
os.chdir(work)
os.environ["DIFFMM_CPU_RNG"] = "1"
out = {}
for prec, ft, fl in itertools.product(("bf16x3", "bf16"), ("1", "0"), ("1", "0")):
    os.environ["DIFFMM_FUSED_TRAIN"], os.environ["DIFFMM_FUSED_LOSS"] = ft, fl
    cfg = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
    cfg.train.epoch = epochs
    cfg.base.precision = prec
    Main.seed_it(cfg.base.seed)
    h = Main.DataHandler(cfg)
    h.LoadData()
    coach = Main.Coach(h, cfg)
    coach.run()
    key = f"{prec} fused_train={ft} fused_loss={fl}"
    out[key] = [dict(train=r["train"], test=r.get("test")) for r in coach.history]
    print(key, " | ".join(f"R@20 {r['test']['Recall']:.5f} N@20 {r['test']['NDCG']:.5f} Loss {r['train']['Loss']:.4f} img {r['train']['image loss']:.4f}"
                          for r in coach.history), flush=True)
    del coach, h
    torch.cuda.empty_cache()
for tag in ("result", "noise_floor_3threads", "noise_floor_1threads", "noise_floor_2threads", "noise_floor_4threads", "noise_floor_6threads"):
    p = os.path.join(GOLD, tag + ".json")
    if os.path.isfile(p):
        g = json.load(open(p))
        print(f"reference {tag} ({g.get('threads')} threads)", " | ".join(
            f"R@20 {e['test']['Recall']:.5f} N@20 {e['test']['NDCG']:.5f} Loss {e['train']['Loss']:.4f} img {e['train']['image loss']:.4f}"
            for e in g["epochs"][:epochs]))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tiktok_real_variants.json"), "w"), indent=1)
