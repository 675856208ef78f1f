"""End-to-end parity: two training epochs + eval of the reference trainer (run on CPU through
oracle/gen_epoch_golden.py, committed under tests/golden/epoch_run/) replayed through diffmm_b200's
Coach on the GPU with the same seeds.  DIFFMM_CPU_RNG=1 makes the noise draws come from the CPU
generator in the reference's program order, so the two runs see identical random numbers.

Tolerance: per-epoch losses rel 2e-3 (fp32-faithful bf16x3 contractions; the rebuilt graphs may differ
in a few near-tie edges), Recall@20 / NDCG@20 within 0.5 % absolute of the metric scale 1 (north_star)."""
import json
import os
import shutil

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "epoch_run")


def _run(tmp_path, precision, monkeypatch, cpu_rng=True, cuda_graph=False):
    from diffmm_b200 import Main
    from diffmm_b200.Conf import Config
    gold = json.load(open(os.path.join(GOLD, "result.json")))
    shutil.copytree(os.path.join(GOLD, "Datasets"), tmp_path / "Datasets")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DIFFMM_CPU_RNG", "1" if cpu_rng else "0")
    cfg = Config()
    cfg.data.name = "tiktok"
    for k, v in gold["overrides"].items():
        sec, key = k.split(".")
        setattr(getattr(cfg, sec), key, v)
    cfg.base.precision = precision
    cfg.base.cuda_graph = cuda_graph
    Main.seed_it(cfg.base.seed)
    handler = Main.DataHandler(cfg)
    handler.LoadData()
    coach = Main.Coach(handler, cfg)
    coach.run()
    return gold, coach


def test_two_epochs_match_reference_bf16x3(tmp_path, monkeypatch):
    gold, coach = _run(tmp_path, "bf16x3", monkeypatch)
    assert len(coach.history) == len(gold["epochs"]) == 2
    for got, want in zip(coach.history, gold["epochs"]):
        for k, v in want["train"].items():
            assert got["train"][k] == pytest.approx(v, rel=2e-3), (k, got["train"][k], v)
        for k in ("Recall", "NDCG", "Precision"):
            assert abs(got["test"][k] - want["test"][k]) <= 0.005, (k, got["test"][k], want["test"][k])


def test_two_epochs_bf16_within_tolerance(tmp_path, monkeypatch):
    gold, coach = _run(tmp_path, "bf16", monkeypatch)
    for got, want in zip(coach.history, gold["epochs"]):
        for k in ("Loss", "BPR Loss", "reg loss", "CL loss"):
            assert got["train"][k] == pytest.approx(want["train"][k], rel=2e-2), (k, got["train"][k], want["train"][k])
        assert abs(got["test"]["Recall"] - want["test"]["Recall"]) <= 0.02


def test_device_rng_run_is_sane(tmp_path, monkeypatch):
    # default mode (device generator, like the reference on a GPU): statistical agreement only
    gold, coach = _run(tmp_path, "bf16", monkeypatch, cpu_rng=False)
    for got, want in zip(coach.history, gold["epochs"]):
        assert got["train"]["Loss"] == pytest.approx(want["train"]["Loss"], rel=0.1)
        assert torch.isfinite(torch.tensor(list(got["train"].values()))).all()


def test_cuda_graph_joint_training_tracks_eager(tmp_path, monkeypatch):
    """Phase 3 replayed from a CUDA graph (base.cuda_graph) vs the eager loop, device generator, same seeds: the
    first epoch sees the same batches, negatives and (graph-safe Philox) noise, so its losses agree closely;
    capturable Adam and the graph's kernel order only move low-order bits."""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    _, eager = _run(tmp_path / "a", "bf16", monkeypatch, cpu_rng=False, cuda_graph=False)
    _, graph = _run(tmp_path / "b", "bf16", monkeypatch, cpu_rng=False, cuda_graph=True)
    assert graph._use_graph() and not eager._use_graph()
    e0, g0 = eager.history[0]["train"], graph.history[0]["train"]
    for k in ("Loss", "BPR Loss", "reg loss", "CL loss"):
        assert g0[k] == pytest.approx(e0[k], rel=5e-3), (k, g0[k], e0[k])
    # phase 1 is graph-replayed too (after two eager warm-up batches): same batches and timesteps, graph-safe noise draws
    assert graph._diff_graph is not None or len(graph.handler.diffusionLoader) <= 3
    for k in ("image loss", "text loss", "audio loss"):
        assert g0[k] == pytest.approx(e0[k], rel=0.2), (k, g0[k], e0[k])
    e1, g1 = eager.history[1]["train"], graph.history[1]["train"]
    assert g1["Loss"] == pytest.approx(e1["Loss"], rel=5e-2)
    assert abs(graph.history[1]["test"]["Recall"] - eager.history[1]["test"]["Recall"]) <= 0.03


def test_device_eval_equals_host_eval(tmp_path, monkeypatch):
    """testEpoch (mask + top-K + metrics kernels, one host sync) vs testEpochHost (torch.topk + the reference's calcRes
    arithmetic on the host): same users, same scores -> identical Recall / Precision, NDCG to the last bits."""
    _, coach = _run(tmp_path, "bf16x3", monkeypatch)
    dev = coach.testEpoch()
    host = coach.testEpochHost()
    assert dev["Recall"] == host["Recall"] and dev["Precision"] == host["Precision"]
    assert dev["NDCG"] == pytest.approx(host["NDCG"], rel=1e-13)
    assert dev["Recall"] > 0


def test_cuda_graph_mode_over_epochs_keeps_derived_state_fresh(tmp_path, monkeypatch):
    """graph replay updates the weights without bumping their version counters, so everything derived from
    them (packed operand copies, hidden-space operators, static adjacency buffers + SpMM plans) relies on explicit
    invalidation.  Dataset with 5 full diffusion batches + a ragged tail and > 5 full joint batches, 3 epochs: both graphs
    must be captured, and after the last epoch the adjacencies the trainer holds must equal a from-scratch eager rebuild
    with the trainer's current weights, and the cached packs must equal fresh packs of the current weights."""
    from diffmm_b200 import Main, autograd as ag, ops, rebuild, synth
    from diffmm_b200.Conf import Config
    U, I = 700, 300
    synth.write_dataset(str(tmp_path), "tiktok", synth.interactions(U, I, seed=3, mean_deg=6.0, heavy_frac=0.02),
                        synth.features(I, dict(image=16, text=24, audio=8), seed=3))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DIFFMM_CPU_RNG", "0")
    cfg = Config()
    cfg.data.name = "tiktok"
    cfg.base.denoise_dim, cfg.base.cuda_graph, cfg.base.precision = "[64]", True, "bf16"
    cfg.train.batch, cfg.train.test_batch, cfg.train.epoch = 128, 128, 3
    Main.seed_it(7)
    handler = Main.DataHandler(cfg)
    handler.LoadData()
    coach = Main.Coach(handler, cfg)
    assert len(handler.diffusionLoader) == 6 and U % 128 != 0 and len(handler.trainLoader) > 6
    coach.run()
    assert coach._use_graph() and coach._diff_graph is not None and coach._joint_graph is not None
    assert all(np.isfinite(list(h["train"].values())).all() for h in coach.history)
    # the adjacencies held after the last epoch vs an eager rebuild from the same (current) weights.  The last epoch's
    # rebuild ran before its joint phase changed nothing the rebuild depends on (Denoise weights move only in phase 1).
    ag._PACK_CACHE.clear()
    for den in coach._denoise_dict().values():
        den._dmm_hidden_ops = None
    fresh = rebuild.rebuild_modal_adj(coach.diffusion_model, coach._denoise_dict(), handler.train_indptr, handler.train_indices,
                                      U, I, cfg.hyper.sampling_step, "bf16")
    for name, adj in (("image", coach.image_adj), ("text", coach.text_adj), ("audio", coach.audio_adj)):
        assert torch.equal(adj.ptr, fresh[name].ptr) and torch.equal(adj.idx, fresh[name].idx), name
        assert torch.equal(adj.val, fresh[name].val), name
        x = torch.randn((U + I, 64), device=adj.val.device)
        assert torch.equal(ops.spmm(adj, x), ops.spmm(fresh[name], x)), name      # the in-place re-planned SpMM plan too
    # cached packs of the Denoise weights == fresh packs of the current values
    for den in coach._denoise_dict().values():
        w = den.in_layers[0].weight
        hi_cached, _ = ag.packed_weight(w, False, False)
        hi_fresh, _ = ops.pack_bf16(w.detach(), split=False)
        assert torch.equal(hi_cached, hi_fresh)
