"""DataHandler with the reference's interface (reference DataHandler.py:10-228).

Same file formats (scipy COO pickles, .npy features under ./Datasets/<name>/), same attributes
(trainMat, torchBiAdj, trainData/trainLoader, testData/testLoader, *_feats, diffusionData/
diffusionLoader, user_pos_items) and the same RNG consumption (numpy negatives, torch CPU-generator
shuffles).  What changes underneath:
  * the train matrix lives on the device as CSR (indptr int64, indices int32); dense user rows are
    materialised per batch by dmm_csr_rows_to_dense instead of keeping the dense U x I matrix
    (DataHandler.py:128: 548 MB at baby, 4 TB at 2M x 500k);
  * makeTorchAdj builds the normalised adjacency on the device (dmm_build_norm_adj_csr) and returns a
    real torch sparse COO tensor (so an unchanged Main.py can still call torch.sparse.mm on it) that
    carries its CSR twin as ``_dmm_csr`` for the SpMM kernel;
  * any dataset directory name is accepted; audio features are used when audio_feat.npy exists.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import scipy.sparse as sp
import torch
import torch.utils.data as dataloader
from scipy.sparse import coo_matrix, csr_matrix
from torch.utils.data import Dataset as torch_dataset

from . import ops


def csr_arrays_from_scipy(mat) -> tuple:
    """Binary, duplicate-free, column-sorted CSR arrays of a scipy matrix (host)."""
    csr = csr_matrix(mat)
    csr.sum_duplicates()
    csr.sort_indices()
    csr.eliminate_zeros()
    return csr.indptr.astype(np.int64), csr.indices.astype(np.int32)


def csr_from_torch_sparse(adj: torch.Tensor) -> ops.CsrAdj:
    """CSR twin of an arbitrary torch sparse N x N tensor (values kept as they are)."""
    a = adj.coalesce()
    idx, val = a.indices(), a.values().float()
    N = a.shape[0]
    counts = torch.bincount(idx[0], minlength=N)
    ptr = torch.zeros(N + 1, dtype=torch.int64, device=idx.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ops.CsrAdj(ptr, idx[1].to(torch.int32).contiguous(), val.contiguous(), N, 0)


class DataHandler:
    def __init__(self, config):
        self.config = config
        self.device = torch.device(f"cuda:{self.config.base.gpu}" if torch.cuda.is_available() else "cpu")
        predir = os.path.join(".", "Datasets", str(self.config.data.name)) + "/"
        if not os.path.isdir(predir):
            raise ValueError(f"Unknown dataset: {self.config.data.name}")
        self.predir = predir
        self.trainfile = predir + "trnMat.pkl"
        self.testfile = predir + "tstMat.pkl"
        self.imagefile = predir + "image_feat.npy"
        self.textfile = predir + "text_feat.npy"
        self.audiofile = predir + "audio_feat.npy"
        self.has_audio = os.path.isfile(self.audiofile)      # the reference keys this on name == 'tiktok'

    def loadOneFile(self, filename):
        """DataHandler.py:41-50."""
        with open(filename, "rb") as fs:
            ret = (pickle.load(fs) != 0).astype(np.float32)
        if not isinstance(ret, coo_matrix):
            ret = coo_matrix(ret)
        return ret

    @staticmethod
    def normalizeAdj(mat: coo_matrix):
        """Host restatement kept for API parity (DataHandler.py:53-66); the hot path uses the device build."""
        csr_mat = mat.tocsr()
        degree = np.asarray(csr_mat.sum(axis=1)).squeeze()
        dInvSqrt = np.where(degree > 0, degree ** (-0.5), 0)
        dInvSqrtMat = sp.diags(dInvSqrt, offsets=0, format="csr")
        return (dInvSqrtMat @ mat @ dInvSqrtMat).tocoo()

    @staticmethod
    def makeCsrAdj(mat, rows: int, cols: int, device) -> ops.CsrAdj:
        indptr, indices = csr_arrays_from_scipy(mat)
        return ops.build_norm_adj(torch.from_numpy(indptr).to(device), torch.from_numpy(indices).to(device), rows, cols)

    @staticmethod
    def makeTorchAdj(mat: coo_matrix, rows: int, cols: int, device):
        """DataHandler.py:69-93: normalised bipartite adjacency as a torch sparse (N, N) tensor."""
        csr = DataHandler.makeCsrAdj(mat, rows, cols, device)
        t = csr.to_torch_coo()
        t._dmm_csr = csr
        return t

    def loadFeatures(self, filename):
        feats = np.load(filename)
        return torch.tensor(feats, dtype=torch.float, device=self.device), feats.shape[1]

    def LoadData(self):
        """DataHandler.py:107-131."""
        trainMat = self.loadOneFile(self.trainfile)
        testMat = self.loadOneFile(self.testfile)
        self.trainMat = trainMat
        self.config.data.user_num, self.config.data.item_num = trainMat.get_shape()
        U, I = self.config.data.user_num, self.config.data.item_num
        indptr, indices = csr_arrays_from_scipy(trainMat)
        self.train_indptr = torch.from_numpy(indptr).to(self.device)
        self.train_indices = torch.from_numpy(indices).to(self.device)
        self.torchBiAdj = self.makeTorchAdj(trainMat, U, I, self.device)

        self.trainData = TrainData(trainMat, self.config)
        self.trainLoader = TrainLoader(self.trainData, self.config.train.batch)
        self.testData = TestData(testMat, trainMat)
        self.testLoader = dataloader.DataLoader(self.testData, batch_size=self.config.train.test_batch, shuffle=False, num_workers=0)

        self.image_feats, self.config.data.image_feat_dim = self.loadFeatures(self.imagefile)
        self.text_feats, self.config.data.text_feat_dim = self.loadFeatures(self.textfile)
        if self.has_audio:
            self.audio_feats, self.config.data.audio_feat_dim = self.loadFeatures(self.audiofile)

        self.diffusionData = DiffusionData(self.train_indptr, self.train_indices, U, I, self.config)
        self.diffusionLoader = DiffusionLoader(self.diffusionData, self.config.train.batch)
        self.user_pos_items = self.trainData.user_pos_items

    def getUserDegrees(self) -> np.ndarray:
        """DataHandler.py:133-143."""
        if not hasattr(self, "trainMat"):
            raise ValueError("Training matrix not loaded. Please call LoadData() first.")
        return np.asarray(self.trainMat.sum(axis=1), dtype=int).squeeze()


class TrainData(torch_dataset):
    """DataHandler.py:145-179 (same negative sampler, same numpy RNG stream)."""

    def __init__(self, coomat: coo_matrix, config):
        self.config = config
        self.rows = coomat.row
        self.cols = coomat.col
        self.dokmat = coomat.todok()
        self.negs = np.zeros(len(self.rows)).astype(np.int32)
        self.user_pos_items = [[] for _ in range(coomat.get_shape()[0])]
        for u, i in zip(self.rows, self.cols):
            self.user_pos_items[u].append(i)
        self._keys = None
        self._csr = None

    def negSamplingLoop(self):
        """The reference's loop verbatim in behaviour (DataHandler.py:159-169); kept as the checker of negSampling."""
        item_num = self.config.data.item_num
        dok = self.dokmat
        rows = self.rows
        for i in range(len(rows)):
            u = rows[i]
            while True:
                neg_index = np.random.randint(item_num)
                if (u, neg_index) not in dok:
                    break
            self.negs[i] = neg_index

    def negSampling(self):
        """One rejection-sampled negative per interaction, in COO order (DataHandler.py:159-169), WITHOUT changing
        the numpy RNG stream: legacy ``np.random.randint(n, size=T)`` yields exactly the values of T scalar calls,
        so a block of draws is taken up front and the reference's loop is replayed over it in native code
        (dmm_host_neg_sampling: interaction i consumes draws until one is not an item of its user).  The global
        generator ends in the state the loop would leave: it is rewound and advanced by the draws consumed."""
        import ctypes as C
        from . import _lib
        item_num = int(self.config.data.item_num)
        n = len(self.rows)
        if n == 0:
            return
        if self._csr is None:
            self._csr = csr_arrays_from_scipy(coo_matrix((np.ones(n), (self.rows, self.cols)),
                                                         shape=(int(self.rows.max()) + 1, item_num)))
        indptr, indices = self._csr
        rows = np.ascontiguousarray(self.rows, dtype=np.int32)
        negs = np.empty(n, dtype=np.int32)
        consumed = C.c_int64(0)
        state = np.random.get_state()
        T = n + max(4096, n // 16)
        lib = _lib.load()
        while True:
            np.random.set_state(state)
            draws = np.ascontiguousarray(np.random.randint(item_num, size=T), dtype=np.int64)
            rc = lib.dmm_host_neg_sampling(indptr.ctypes.data, indices.ctypes.data, rows.ctypes.data, n,
                                           draws.ctypes.data, T, negs.ctypes.data, C.byref(consumed))
            if rc == 0:
                break
            if rc != -4:                      # DMM_ERR_WORKSPACE: the stream ran out, draw a longer one
                _lib.check(rc, "dmm_host_neg_sampling")
            T *= 2
        np.random.set_state(state)
        np.random.randint(item_num, size=int(consumed.value))
        self.negs = negs

    def negSamplingFast(self, rng: np.random.Generator):
        """Vectorised variant (different RNG stream: statistical parity only; opt-in)."""
        item_num = self.config.data.item_num
        if self._keys is None:
            self._keys = np.unique(self.rows.astype(np.int64) * item_num + self.cols.astype(np.int64))
        negs = rng.integers(0, item_num, len(self.rows))
        todo = np.arange(len(self.rows))
        while len(todo):
            bad = np.isin(self.rows[todo].astype(np.int64) * item_num + negs[todo], self._keys)
            todo = todo[bad]
            negs[todo] = rng.integers(0, item_num, len(todo))
        self.negs = negs.astype(np.int32)

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, idx):
        return self.rows[idx], self.cols[idx], self.negs[idx]


class TrainLoader:
    """Iterates (users, pos, neg) batches like DataLoader(TrainData, batch, shuffle=True) (DataHandler.py:117-118).
    The index stream comes from a real torch DataLoader over range(E), so the CPU-generator consumption (base
    seed, sampler seed, randperm) is the reference's; the per-sample __getitem__ + default_collate of 1024
    numpy scalars per batch is replaced by three vectorised gathers (same values, same int32 dtype)."""

    def __init__(self, data: "TrainData", batch_size: int):
        self.dataset = data
        self.batch_size = batch_size
        self._index_loader = dataloader.DataLoader(_IndexOnly(len(data)), batch_size=batch_size, shuffle=True, num_workers=0)

    def __len__(self):
        return len(self._index_loader)

    def __iter__(self):
        d = self.dataset
        for idx in self._index_loader:
            i = idx.numpy()
            yield torch.from_numpy(d.rows[i]), torch.from_numpy(d.cols[i]), torch.from_numpy(np.asarray(d.negs)[i])


class TestData(torch_dataset):
    """DataHandler.py:181-209."""

    def __init__(self, testMat: coo_matrix, trainMat: coo_matrix):
        self.trainMat_csr = sp.csr_matrix(trainMat.tocsr() != 0) * 1.0
        test_use_its = [None] * testMat.get_shape()[0]
        test_users = set()
        for i in range(len(testMat.data)):
            user_idx = testMat.row[i]
            item_idx = testMat.col[i]
            if test_use_its[user_idx] is None:
                test_use_its[user_idx] = list()
            test_use_its[user_idx].append(item_idx)
            test_users.add(user_idx)
        self.test_users = np.array(list(test_users))
        self.test_user_its = test_use_its

    def __len__(self):
        return len(self.test_users)

    def __getitem__(self, idx):
        return self.test_users[idx], np.reshape(self.trainMat_csr[self.test_users[idx]].toarray(), [-1])


class DiffusionData(torch_dataset):
    """CSR-resident replacement of the dense U x I tensor (DataHandler.py:211-228)."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, n_users: int, n_items: int, config):
        self.indptr, self.indices = indptr, indices
        self.n_users, self.n_items = n_users, n_items
        self.device = indptr.device

    def rows(self, row_ids: torch.Tensor) -> torch.Tensor:
        """Dense fp32 [len(row_ids), I] rows of the binary train matrix."""
        row_ids = row_ids.to(self.device, torch.int64)
        n = row_ids.numel()
        ld = ops.pad_to(self.n_items, 4)
        out = torch.empty((n, ld), dtype=torch.float32, device=self.device)
        ops.csr_rows_to_dense(self.indptr, self.indices, n, self.n_items, row_ids=row_ids, x_f32=out)
        return out[:, :self.n_items] if ld != self.n_items else out

    def __getitem__(self, index):
        return self.rows(torch.tensor([int(index)]))[0], index

    def __len__(self):
        return self.n_users


class _IndexOnly(torch_dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class DiffusionLoader:
    """Iterates (dense_rows [B, I], user_idx [B]) like DataLoader(DiffusionData, batch, shuffle=True).
    The index stream comes from a real torch DataLoader over range(U), so the CPU-generator consumption
    (base seed, sampler seed, randperm) is the reference's; only the row gather runs on our kernel."""

    def __init__(self, data: DiffusionData, batch_size: int):
        self.dataset = data
        self.batch_size = batch_size
        self._index_loader = dataloader.DataLoader(_IndexOnly(len(data)), batch_size=batch_size, shuffle=True, num_workers=0)

    def __len__(self):
        return len(self._index_loader)

    def index_batches(self):
        yield from self._index_loader

    def __iter__(self):
        for idx in self._index_loader:
            yield self.dataset.rows(idx), idx
