"""Chaos floor of the real-TikTok epoch run under OUR arithmetic: the same run (tests/golden/tiktok_real, reference CPU
RNG stream replayed) repeated with the initial Denoise weights moved by +-1 ulp at random (a perturbation of 6e-8
relative: far below any rounding the reference itself commits).  The spread of Recall@20 / NDCG@20 / losses over the
members is what a bit-level change of the arithmetic does to the trajectory; a systematic error of a code path shows
as a shift of its ensemble MEAN against the per-op fp32-faithful ensemble and against the reference's runs.
    python tools/tiktok_real_ensemble.py [members] [epochs] [out.json]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, ".")
ROOT = os.path.abspath(".")
GOLD = os.path.join(ROOT, "tests", "golden", "tiktok_real")
members = int(sys.argv[1]) if len(sys.argv) > 1 else 5
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "tiktok_real_ensemble.json")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle.gen_tiktok_golden import materialise  # noqa: E402
from diffmm_b200 import Main  # noqa: E402
from diffmm_b200.Conf import load_config  # noqa: E402

work = tempfile.mkdtemp(prefix="tiktok_ens_")
materialise(GOLD, work)
os.chdir(work)
os.environ["DIFFMM_CPU_RNG"] = "1"


class PerturbedCoach(Main.Coach):
    member = 0

    def prepareModel(self):
        super().prepareModel()
        if self.member == 0:
            return
        rs = np.random.default_rng(1000 + self.member)          # not torch's generator: the run's RNG stream is untouched
        with torch.no_grad():
            for den in self._denoise_dict().values():
                for p in den.parameters():
                    step = torch.from_numpy(rs.integers(-1, 2, size=tuple(p.shape)).astype(np.int8)).to(p.device)
                    up = torch.nextafter(p, torch.full_like(p, float("inf")))
                    down = torch.nextafter(p, torch.full_like(p, float("-inf")))
                    p.copy_(torch.where(step > 0, up, torch.where(step < 0, down, p)))


def run(prec, fused_train, member):
    os.environ["DIFFMM_FUSED_TRAIN"] = fused_train
    cfg = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
    cfg.train.epoch = epochs
    cfg.base.precision = prec
    Main.seed_it(cfg.base.seed)
    h = Main.DataHandler(cfg)
    h.LoadData()
    coach = PerturbedCoach(h, cfg)
    coach.member = member
    coach.run()
    hist = [dict(train=r["train"], test=r.get("test")) for r in coach.history]
    del coach, h
    torch.cuda.empty_cache()
    return hist


out = {}
for prec, ft in (("bf16x3", "0"), ("bf16x3", "1"), ("bf16", "0"), ("bf16", "1")):
    key = f"{prec} fused_train={ft}"
    out[key] = []
    for m in range(members):
        hist = run(prec, ft, m)
        out[key].append(hist)
        print(key, f"member {m}", " | ".join(f"R@20 {r['test']['Recall']:.5f} N@20 {r['test']['NDCG']:.5f} Loss {r['train']['Loss']:.4f}"
                                             for r in hist), flush=True)
    for e in range(epochs):
        rec = np.array([h[e]["test"]["Recall"] for h in out[key]])
        nd = np.array([h[e]["test"]["NDCG"] for h in out[key]])
        print(f"   epoch {e}: Recall@20 mean {rec.mean():.5f} sd {rec.std(ddof=1) if members > 1 else 0:.5f} min {rec.min():.5f} max {rec.max():.5f}"
              f" | NDCG@20 mean {nd.mean():.5f} sd {nd.std(ddof=1) if members > 1 else 0:.5f}", flush=True)
    json.dump(out, open(out_path, "w"), indent=1)
for tag in ("result", "noise_floor_1threads", "noise_floor_2threads", "noise_floor_3threads", "noise_floor_4threads",
            "noise_floor_6threads"):
    p = os.path.join(GOLD, tag + ".json")
    if os.path.isfile(p):
        g = json.load(open(p))
        print(f"reference {tag} ({g.get('threads')} threads)", " | ".join(
            f"R@20 {e['test']['Recall']:.5f} N@20 {e['test']['NDCG']:.5f} Loss {e['train']['Loss']:.4f}" for e in g["epochs"][:epochs]))
