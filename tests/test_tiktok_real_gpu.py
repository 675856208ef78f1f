"""Real-config parity gate (BASELINE.json configs[0]): conf/tiktok.toml on the REAL TikTok
interactions + image / audio features (9308 x 6710, hidden 1024, 3 modalities), three epochs + eval.

Golden: the UNMODIFIED reference run on CPU by oracle/gen_tiktok_golden.py -> tests/golden/tiktok_real/result.json (8 torch
threads; reproduces BASELINE.md's epoch-0 log: Loss 4.37246, Recall@20 0.05546).

What "equal to the reference" can mean here was measured, not assumed.  The run is chaotic in its ranking metrics: Adam's
first steps move every weight by +-lr whatever the size of its gradient, so a bit-level change of the arithmetic changes
the trajectory.  tests/golden/tiktok_real holds 14 runs of the reference itself that differ only at that level --
result.json, five other torch thread counts (noise_floor_*threads.json: another fp32 summation order), and eight runs
whose initial Denoise weights were moved by -1/0/+1 ulp at random (ensemble_perturb*.json).  Their Recall@20 is
0.05430 +- 0.00102 / 0.06733 +- 0.00134 / 0.07235 +- 0.00106 (mean +- sd over the 14 runs, epochs 0 / 1 / 2): the
reference's own run-to-run sd is 1.5-2.0 %, four times north_star's 0.5 %.  The smooth epoch losses move by 0.05-0.17 %.
So the gates are:
  * smooth epoch losses (Loss / BPR / reg / CL): |ours - ref mean| <= 0.5 %, every epoch, every run of ours;
  * Recall / NDCG / Precision @20 of ONE run of ours: within 3.5 sd of the reference ensemble mean, sd = the larger of
    the reference's own sd and 1.5 % (the sd of OUR ensembles: our runs are chaotic too, and not bit-reproducible --
    the loss backward scatters with atomics), and epoch 0 / 2 near BASELINE.md's published 0.05546 / 0.07341 at that level;
  * Recall@20 / NDCG@20 of the benchmarked precision (bf16) as an ENSEMBLE MEAN over 6 members of the same perturbation
    (tools/tiktok_real_ensemble.py) against the reference ensemble mean: within 3 % and within 4.5 standard errors of
    the difference -- the test that can see a systematic shift, which a single chaotic run cannot.  Measured offsets of
    our ensemble means: -0.8 % .. -1.4 % (1.3 .. 2 standard errors): compatible with zero, a shift of ~1 % cannot be
    excluded with ensembles of this size;
  * the logged per-modality diffusion "losses" (a running quantity renormalised every batch, Main.py:177-185, dominated
    by the 92-user tail batch with SNR weights up to 9.6e3; reference sd 2-17 %): <= max(5 %, 4 sd).
Measured in round 2 (profiles/r02_tiktok_real_ensemble.txt): our four arithmetic variants (bf16 / bf16x3, fused step /
per-op autograd) give ensemble means 0.0534-0.0539 +- 0.0008 at epoch 0 -- indistinguishable from each other and within
2 standard errors of the reference's 0.0543.
Our runs replay the reference's CPU RNG stream (DIFFMM_CPU_RNG=1)."""
import glob
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "tiktok_real")
SMOOTH = ("Loss", "BPR Loss", "reg loss", "CL loss")
RANKING = ("Recall", "NDCG", "Precision")


def _reference_ensemble():
    """{(epoch, section, key): (mean, sd, n)} over the committed reference runs."""
    runs = [json.load(open(p)) for p in sorted(glob.glob(os.path.join(GOLD, "*.json")))]
    assert len(runs) >= 14
    stats = {}
    for e in range(3):
        for sec in ("train", "test"):
            for k in runs[0]["epochs"][e][sec]:
                v = np.array([r["epochs"][e][sec][k] for r in runs])
                stats[(e, sec, k)] = (float(v.mean()), float(v.std(ddof=1)), len(v))
    return stats


def _run(tmp_path, monkeypatch, precision, epochs, member=0):
    sys.path.insert(0, ROOT)
    import torch
    from oracle.gen_tiktok_golden import materialise      # dataset writer only (numpy / scipy; no reference import)
    from diffmm_b200 import Main
    from diffmm_b200.Conf import load_config
    if not os.path.isdir(os.path.join(str(tmp_path), "Datasets")):
        materialise(GOLD, str(tmp_path))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DIFFMM_CPU_RNG", "1")

    class PerturbedCoach(Main.Coach):
        def prepareModel(self):
            super().prepareModel()
            if member == 0:
                return
            rs = np.random.default_rng(1000 + member)      # the same +-1 ulp moves as oracle/gen_tiktok_golden.py --perturb
            with torch.no_grad():
                for den in self._denoise_dict().values():
                    for p in den.parameters():
                        step = torch.from_numpy(rs.integers(-1, 2, size=tuple(p.shape)).astype(np.int8)).to(p.device)
                        up = torch.nextafter(p, torch.full_like(p, float("inf")))
                        down = torch.nextafter(p, torch.full_like(p, float("-inf")))
                        p.copy_(torch.where(step > 0, up, torch.where(step < 0, down, p)))

    cfg = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
    cfg.train.epoch = epochs
    cfg.base.precision = precision
    Main.seed_it(cfg.base.seed)
    handler = Main.DataHandler(cfg)
    handler.LoadData()
    coach = PerturbedCoach(handler, cfg)
    coach.run()
    hist = coach.history
    del coach, handler
    torch.cuda.empty_cache()
    return hist


def _rows(hist, stats, name):
    rows = []
    for e, got in enumerate(hist):
        for sec in ("train", "test"):
            for k, v in got[sec].items():
                mu, sd, n = stats[(e, sec, k)]
                sd_gate = max(sd, 0.015 * abs(mu))
                rows.append(dict(epoch=e, key=k, ours=v, ref_mean=mu, ref_sd=sd, ref_runs=n, rel_err=abs(v - mu) / abs(mu),
                                 z=(v - mu) / sd if sd > 0 else 0.0, z_gate=(v - mu) / sd_gate if sd_gate > 0 else 0.0))
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(rows, open(os.path.join(out, f"tiktok_real_parity_{name}.json"), "w"), indent=1)
    except OSError:
        pass
    return rows


def _check_single_run(rows):
    bad = []
    for r in rows:
        if r["key"] in SMOOTH:
            ok = r["rel_err"] <= 0.005
        elif r["key"] in RANKING:
            ok = abs(r["z_gate"]) <= 3.5
        else:
            ok = abs(r["ours"] - r["ref_mean"]) <= max(0.05 * abs(r["ref_mean"]), 4.0 * r["ref_sd"])
        if not ok:
            bad.append(r)
    assert not bad, bad


def test_tiktok_real_three_epochs_bf16x3(tmp_path, monkeypatch):
    stats = _reference_ensemble()
    hist = _run(tmp_path, monkeypatch, "bf16x3", 3)
    _check_single_run(_rows(hist, stats, "bf16x3"))
    # BASELINE.md section 2's published numbers, for the record (sd of the reference's own runs: 0.00102 / 0.00106)
    assert hist[0]["train"]["Loss"] == pytest.approx(4.37246, rel=5e-3)
    assert abs(hist[0]["test"]["Recall"] - 0.05546) <= 3.5 * stats[(0, "test", "Recall")][1] + 0.00116     # 0.05546 is itself +1.1 sd
    assert abs(hist[2]["test"]["Recall"] - 0.07341) <= 3.5 * stats[(2, "test", "Recall")][1] + 0.00106


def test_tiktok_real_bf16_ensemble_mean(tmp_path, monkeypatch):
    """The benchmarked precision (single-pass bf16 contractions, bf16 propagation table): six members of the +-1 ulp
    ensemble, one epoch each.  Every member: smooth losses within 0.5 % and ranking metrics within 3.5 sd of the reference
    ensemble (single-run gate of the module docstring); the ensemble MEAN of Recall@20 / NDCG@20 within 3 % and 4.5 standard
    errors of the reference ensemble mean."""
    stats = _reference_ensemble()
    monkeypatch.setenv("DIFFMM_SPMM_BF16_MIN_NNZ", "0")      # the bf16 propagation table too (by default only from 1 M entries)
    members = 6
    hists = [_run(tmp_path, monkeypatch, "bf16", 1, member=m) for m in range(members)]
    for m, h in enumerate(hists):
        _check_single_run(_rows(h, stats, f"bf16_member{m}"))
    report = {}
    for k in ("Recall", "NDCG"):
        ours = np.array([h[0]["test"][k] for h in hists])
        mu, sd, n = stats[(0, "test", k)]
        se = float(np.sqrt(sd * sd / n + ours.var(ddof=1) / members))
        report[k] = dict(ours_mean=float(ours.mean()), ours_sd=float(ours.std(ddof=1)), ref_mean=mu, ref_sd=sd, ref_runs=n,
                         std_err_of_difference=se, z=float((ours.mean() - mu) / se), rel_diff=float(ours.mean() / mu - 1.0))
    try:
        json.dump(report, open(os.path.join(ROOT, "gpurun_out", "tiktok_real_parity_bf16_ensemble.json"), "w"), indent=1)
    except OSError:
        pass
    for k, r in report.items():
        assert abs(r["z"]) <= 4.5, (k, r)
        assert abs(r["rel_diff"]) <= 0.03, (k, r)
