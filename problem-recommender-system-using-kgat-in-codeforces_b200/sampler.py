"""Vectorised host-side batch samplers with the reference's sampling *semantics*
(src/model/KGAT/preprocess.py:328-530): a CF batch is ``B`` distinct users (with replacement only if
there are fewer than ``B``), one uniformly drawn positive item per user and one uniformly drawn item
the user has not interacted with; a KG batch is ``B`` distinct heads, one uniformly drawn
``(relation, tail)`` of the head and one uniformly drawn node that is not a tail of ``(head,
relation)``.  The reference draws them with per-sample Python loops on an unseeded module-level
Generator, so parity is distributional; the RNG-stream-exact replay lives in the
oracle (``oracle/kgat_oracle.py: sample_cf_batch / sample_kg_batch``) and is tested there.

Used by the benchmark / epoch driver to pre-sample an epoch's batches (the samplers are outside
the measured hot path, which starts at ``model(...)``).
"""

from __future__ import annotations

import numpy as np

from .ckg import CKG


class BatchSampler:
    def __init__(self, g: CKG, seed: int = 2024):
        self.g = g
        self.rng = np.random.default_rng(seed)
        tr = g.train_interactions
        order = np.lexsort((tr[:, 1], tr[:, 0]))
        tr = tr[order]
        self._items = tr[:, 1].astype(np.int64)
        counts = np.bincount(tr[:, 0], minlength=g.user_num)
        self._uptr = np.concatenate([[0], np.cumsum(counts)])
        self._users = np.nonzero(counts)[0].astype(np.int64)
        self._ui_keys = tr[:, 0].astype(np.int64) * g.item_num + self._items  # sorted
        n = g.node_num
        hcounts = np.bincount(g.heads, minlength=n)
        self._hptr = np.concatenate([[0], np.cumsum(hcounts)])
        self._heads = np.nonzero(hcounts)[0].astype(np.int64)
        self._rel_span = int(g.relations.max()) + 1 if g.nnz else 1
        self._hrt_keys = np.sort((g.heads.astype(np.int64) * self._rel_span + g.relations) * n + g.tails)

    # number of batches per epoch, as the reference driver computes them (main.py:297, 324)
    def cf_batches_per_epoch(self, batch: int = 256) -> int:
        return self.g.train_interactions.shape[0] // batch + 1

    def kg_batches_per_epoch(self, batch: int = 512) -> int:
        return self.g.nnz // batch + 1

    @staticmethod
    def _contains(sorted_keys: np.ndarray, keys: np.ndarray) -> np.ndarray:
        pos = np.searchsorted(sorted_keys, keys)
        pos[pos >= sorted_keys.size] = sorted_keys.size - 1
        return sorted_keys[pos] == keys

    def cf_batches(self, n_batches: int, batch: int = 256):
        rng, g = self.rng, self.g
        replace = batch > self._users.size
        users = np.stack([rng.choice(self._users, size=batch, replace=replace) for _ in range(n_batches)])
        deg = self._uptr[users + 1] - self._uptr[users]
        pos = self._items[self._uptr[users] + np.floor(rng.random(users.shape) * deg).astype(np.int64).clip(max=deg - 1)]
        neg = rng.integers(0, g.item_num, size=users.shape)
        for _ in range(64):
            bad = self._contains(self._ui_keys, users * g.item_num + neg)
            if not bad.any():
                break
            neg[bad] = rng.integers(0, g.item_num, size=int(bad.sum()))
        return users, pos, neg

    def kg_batches(self, n_batches: int, batch: int = 512):
        rng, g = self.rng, self.g
        n = g.node_num
        replace = batch > self._heads.size
        heads = np.stack([rng.choice(self._heads, size=batch, replace=replace) for _ in range(n_batches)])
        deg = self._hptr[heads + 1] - self._hptr[heads]
        e = self._hptr[heads] + np.floor(rng.random(heads.shape) * deg).astype(np.int64).clip(max=deg - 1)
        rels = g.relations[e].astype(np.int64)
        pos = g.tails[e].astype(np.int64)
        neg = rng.integers(0, n, size=heads.shape)
        for _ in range(64):
            bad = self._contains(self._hrt_keys, (heads * self._rel_span + rels) * n + neg)
            if not bad.any():
                break
            neg[bad] = rng.integers(0, n, size=int(bad.sum()))
        return heads, rels, pos, neg


class DeviceSampler:
    """Device-resident sampling structures + the sampling kernels (csrc/sampler.cu): user -> sorted items CSR for
    the BPR sampler, head -> (relation, tail) edges sorted by tail for the KG sampler."""

    def __init__(self, g: CKG, device, seed: int = 2024):
        import torch

        from . import ops

        self._ops, self.g, self.seed = ops, g, int(seed)
        tr = g.train_interactions
        order = np.lexsort((tr[:, 1], tr[:, 0]))
        tr = tr[order]
        counts = np.bincount(tr[:, 0], minlength=g.user_num)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).to(device)  # noqa: E731
        self.user_ptr = to(np.concatenate([[0], np.cumsum(counts)]))
        self.user_items = to(tr[:, 1])
        self.active_users = to(np.nonzero(counts)[0])
        hcounts = np.bincount(g.heads, minlength=g.node_num)
        self.head_ptr = to(np.concatenate([[0], np.cumsum(hcounts)]))
        self.edge_rel = to(g.relations)  # edge list is sorted by (head, tail)
        self.edge_tail = to(g.tails)
        self.active_heads = to(np.nonzero(hcounts)[0])

    def cf_batch(self, step_dev, out):
        return self._ops.sample_cf_batch(self.user_ptr, self.user_items, self.active_users, self.g.item_num, self.seed, step_dev, out)

    def kg_batch(self, step_dev, out):
        return self._ops.sample_kg_batch(self.head_ptr, self.edge_rel, self.edge_tail, self.active_heads, self.g.node_num, self.seed + 1,
                                         step_dev, out)
