"""Synthetic datasets with the reference's on-disk layout (Datasets/<name>/{trnMat,tstMat}.pkl scipy COO
pickles with float64 data, {image,text,audio}_feat.npy float32) and in-memory CSR variants for the
benchmarks.  Shapes follow SURVEY.md §8(d): log-normal user degrees (mean ~6.4) with a 1 % heavy tail of
k in [128, 600] like TikTok's bimodal degree distribution, Zipf(~1) item popularity, one held-out test
item per user.  There is no network access, so every benchmark and end-to-end test uses these.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass

import numpy as np
from scipy.sparse import coo_matrix

SHAPES = {
    # name: (users, items, feature dims)   — SURVEY.md §8 header
    "tiktok": (9308, 6710, dict(image=128, text=768, audio=128)),
    "baby": (19445, 7050, dict(image=4096, text=1024)),
    "sports": (35598, 18357, dict(image=4096, text=1024)),
    "ifashion": (300000, 80000, dict(image=64, text=64)),
    "scaleout": (2000000, 500000, dict(image=64, text=64, audio=64)),
}


@dataclass
class SynthInteractions:
    n_users: int
    n_items: int
    indptr: np.ndarray     # int64 [U+1]
    indices: np.ndarray    # int32 [E], ascending and unique inside each row
    test_items: np.ndarray  # int32 [U], -1 when the user has no test item


def user_degrees(rng, n_users, n_items, mean_deg=6.4, heavy_frac=0.01, k_max=600):
    k = np.clip(np.round(rng.lognormal(mean=np.log(mean_deg) - 0.5 * 0.9 ** 2, sigma=0.9, size=n_users)), 1, k_max)
    heavy = rng.random(n_users) < heavy_frac
    k[heavy] = rng.integers(128, k_max + 1, heavy.sum())
    return np.minimum(k.astype(np.int64), max(1, n_items - 1))


def interactions(n_users, n_items, seed=0, mean_deg=6.4, heavy_frac=0.01, zipf_a=1.0) -> SynthInteractions:
    """Vectorised sampler: items drawn from a Zipf-like popularity with replacement, de-duplicated per user
    (so realised degrees are <= the drawn ones), plus a disjoint test item per user."""
    rng = np.random.default_rng(seed)
    k = user_degrees(rng, n_users, n_items, mean_deg, heavy_frac)
    pop = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), zipf_a)
    cdf = np.cumsum(pop / pop.sum())
    perm = rng.permutation(n_items)                       # popularity rank -> item id
    tot = int(k.sum()) + n_users                          # + one test candidate per user
    users = np.repeat(np.arange(n_users, dtype=np.int64), k + 1)
    items = perm[np.minimum(np.searchsorted(cdf, rng.random(tot)), n_items - 1)].astype(np.int64)
    key = np.unique(users * n_items + items)
    users, items = key // n_items, key % n_items
    counts = np.bincount(users, minlength=n_users)
    starts = np.concatenate([[0], np.cumsum(counts)])
    # hold out one interaction of every user that has at least two
    test_items = np.full(n_users, -1, dtype=np.int32)
    pick = starts[:-1] + (rng.random(n_users) * counts).astype(np.int64)
    has = counts >= 2
    test_items[has] = items[pick[has]]
    keep = np.ones(len(users), dtype=bool)
    keep[pick[has]] = False
    users, items = users[keep], items[keep]
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(users, minlength=n_users), out=indptr[1:])
    return SynthInteractions(n_users, n_items, indptr, items.astype(np.int32), test_items)


def features(n_items, dims: dict, seed=0):
    rng = np.random.default_rng(seed + 1)
    return {m: rng.standard_normal((n_items, d), dtype=np.float32) for m, d in dims.items()}


def write_dataset(root: str, name: str, inter: SynthInteractions, feats: dict):
    """Writes ./Datasets/<name>/ under ``root`` in the reference's format (DataHandler.py:18-37)."""
    d = os.path.join(root, "Datasets", name)
    os.makedirs(d, exist_ok=True)
    U, I = inter.n_users, inter.n_items
    rows = np.repeat(np.arange(U), np.diff(inter.indptr))
    trn = coo_matrix((np.ones(len(rows), dtype=np.float64), (rows, inter.indices)), shape=(U, I))
    tu = np.nonzero(inter.test_items >= 0)[0]
    tst = coo_matrix((np.ones(len(tu), dtype=np.float64), (tu, inter.test_items[tu])), shape=(U, I))
    with open(os.path.join(d, "trnMat.pkl"), "wb") as f:
        pickle.dump(trn, f)
    with open(os.path.join(d, "tstMat.pkl"), "wb") as f:
        pickle.dump(tst, f)
    for m, x in feats.items():
        np.save(os.path.join(d, f"{m}_feat.npy"), x)
    return d
