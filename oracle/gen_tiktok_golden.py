"""TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference trainer (Main.Coach.run) on the REAL TikTok interactions
and image / audio features shipped with the reference (Datasets/tiktok, 9308 users x 6710 items, 59541 train
interactions) with the hyper-parameters of conf/tiktok.toml, hidden width 1024, for three epochs + eval on CPU, and
records every epoch's loss dict and Recall/NDCG/Precision@20.  text_feat.npy is not part of the reference checkout
(.MISSING_LARGE_BLOBS), so the text features are ``default_rng(0).standard_normal((6710, 768))`` exactly as in
BASELINE.md section 2, whose numbers (epoch-0 Loss 4.37246, Recall@20 0.05546 -> 0.06949 -> 0.07341) this run reproduces.

The dataset (COO indices as int32, features as stored: float16) and the result are committed under
tests/golden/tiktok_real/ so that the GPU test replays the same run through diffmm_b200 on the GPU box, where
/root/reference does not exist.

    python oracle/gen_tiktok_golden.py [--threads N] [--tag NAME]      # ~2.5 min on 8 cores
"""
import argparse
import json
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "tiktok_real")
EPOCHS = 3


def write_fixture():
    src = os.path.join(ref_shim.REFERENCE_ROOT, "Datasets", "tiktok")
    trn = pickle.load(open(os.path.join(src, "trnMat.pkl"), "rb"))
    tst = pickle.load(open(os.path.join(src, "tstMat.pkl"), "rb"))
    assert (trn.data == 1).all() and (tst.data == 1).all()
    np.savez_compressed(os.path.join(OUT, "dataset.npz"), shape=np.array(trn.shape, dtype=np.int64),
                        trn_row=trn.row.astype(np.int32), trn_col=trn.col.astype(np.int32),
                        tst_row=tst.row.astype(np.int32), tst_col=tst.col.astype(np.int32),
                        image_feat=np.load(os.path.join(src, "image_feat.npy")),
                        audio_feat=np.load(os.path.join(src, "audio_feat.npy")))


def materialise(fixture_dir, dst_root):
    """Datasets/tiktok/ in the reference's on-disk format from the committed fixture (also used by the GPU test)."""
    from scipy.sparse import coo_matrix
    z = np.load(os.path.join(fixture_dir, "dataset.npz"))
    d = os.path.join(dst_root, "Datasets", "tiktok")
    os.makedirs(d, exist_ok=True)
    shape = tuple(int(v) for v in z["shape"])
    for name, r, c in (("trnMat.pkl", z["trn_row"], z["trn_col"]), ("tstMat.pkl", z["tst_row"], z["tst_col"])):
        m = coo_matrix((np.ones(len(r), dtype=np.float64), (r.astype(np.int32), c.astype(np.int32))), shape=shape)
        with open(os.path.join(d, name), "wb") as f:
            pickle.dump(m, f)
    np.save(os.path.join(d, "image_feat.npy"), z["image_feat"])
    np.save(os.path.join(d, "audio_feat.npy"), z["audio_feat"])
    np.save(os.path.join(d, "text_feat.npy"), np.random.default_rng(0).standard_normal((shape[1], 768)).astype(np.float32))
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--tag", default="result")
    ap.add_argument("--perturb", type=int, default=0,
                    help="move every initial Denoise weight by -1/0/+1 ulp at random (numpy seed 1000 + M): the chaos-floor "
                         "ensemble of tools/tiktok_real_ensemble.py on the reference's own arithmetic")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.threads:
        torch.set_num_threads(args.threads)
    if not os.path.isfile(os.path.join(OUT, "dataset.npz")):
        write_fixture()
    work = tempfile.mkdtemp(prefix="tiktok_golden_")
    materialise(OUT, work)
    ref = ref_shim.load_reference()
    cfg = ref.Conf.load_config(os.path.join(ref_shim.REFERENCE_ROOT, "conf", "tiktok.toml"))
    cfg.train.epoch = EPOCHS
    ref.Main.config = cfg
    cwd = os.getcwd()
    os.chdir(work)
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
            print(results[-1], flush=True)
            return r

        def test():
            r = orig_test()
            results[-1]["test"] = {k: float(v) for k, v in r.items()}
            print(results[-1]["test"], flush=True)
            return r

        coach.trainEpoch, coach.testEpoch = train, test
        if args.perturb:
            orig_prepare = coach.prepareModel

            def prepare():
                orig_prepare()
                rs = np.random.default_rng(1000 + args.perturb)
                with torch.no_grad():
                    dens = [coach.image_denoise_model, coach.text_denoise_model, coach.audio_denoise_model]
                    for den in dens:
                        for p in den.parameters():
                            step = torch.from_numpy(rs.integers(-1, 2, size=tuple(p.shape)).astype(np.int8))
                            up = torch.nextafter(p, torch.full_like(p, float("inf")))
                            down = torch.nextafter(p, torch.full_like(p, float("-inf")))
                            p.copy_(torch.where(step > 0, up, torch.where(step < 0, down, p)))

            coach.prepareModel = prepare
        coach.run()
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)
    with open(os.path.join(OUT, f"{args.tag}.json"), "w") as f:
        json.dump({"conf": "conf/tiktok.toml", "epochs_run": EPOCHS, "epochs": results, "torch": torch.__version__,
                   "numpy": np.__version__, "threads": torch.get_num_threads(), "perturb": args.perturb,
                   "text_feat": "np.random.default_rng(0).standard_normal((6710, 768)).astype(float32)"}, f, indent=1)


if __name__ == "__main__":
    main()
