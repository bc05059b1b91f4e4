"""Vectorised host-side batch samplers with the reference's sampling *semantics*
(src/model/KGAT/preprocess.py:328-530): a CF batch is ``B`` distinct users (with replacement only if
there are fewer than ``B``), one uniformly drawn positive item per user and one uniformly drawn item
the user has not interacted with; a KG batch is ``B`` distinct heads, one uniformly drawn
``(relation, tail)`` of the head and one uniformly drawn node that is not a tail of ``(head,
relation)``.  The reference draws them with per-sample Python loops on an unseeded module-level
Generator, so for ``BatchSampler`` / ``DeviceSampler`` parity is distributional.
``ReferenceStreamSampler`` is the bit-exact one: given the same ``numpy.random.Generator`` it returns
the reference's batches id for id (it consumes the generator through the same call sequence; only
the O(degree) ``in list`` membership scans are replaced by hash lookups), tested against batches
recorded from the unmodified ``Preprocess`` under an injected seeded generator.

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


def kg_dict_reference_order(g: CKG):
    """``(heads_in_key_order, ptr, relations, tails)`` of the reference's ``kg_dict`` (preprocess.py:248-266), vectorised.

    The reference walks the Laplacians in ``adjacency_relations`` order and, inside one, scipy's COO of
    ``(D^-1/2 A)^T D^-1/2`` -- a column-major walk of the transposed product, i.e. the merged adjacency entries in
    (adjacency row, adjacency col) order with head = adjacency col.  Dict keys keep first-seen order; a head's list keeps
    walk order.  Both orders feed the samplers (``rng.choice(list(keys))`` and ``positive_triplets[i]``), hence matter
    for stream parity; verified against the recorded reference dict (tests/test_host_logic.py)."""
    n = g.node_num
    rel_pos = {int(r): i for i, r in enumerate(g.adjacency_relations)}
    lap = np.asarray([rel_pos[int(r)] for r in range(max(rel_pos) + 1)], dtype=np.int64)[g.relations]
    # edge (head, rel, tail) of Laplacian k came from adjacency entry (row = tail, col = head): walk order = (k, tail, head)
    order = np.lexsort((g.heads, g.tails, lap))
    heads, rels, tails = g.heads[order].astype(np.int64), g.relations[order].astype(np.int64), g.tails[order].astype(np.int64)
    _, first = np.unique(heads, return_index=True)
    key_heads = heads[np.sort(first)]
    rank = np.empty(n, dtype=np.int64)
    rank[key_heads] = np.arange(key_heads.size)
    by_head = np.argsort(rank[heads], kind="stable")
    counts = np.bincount(rank[heads], minlength=key_heads.size)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    return key_heads, ptr, rels[by_head], tails[by_head]


class ReferenceStreamSampler:
    """Bit-exact replay of ``Preprocess.generate_cf_batch`` / ``generate_kg_batch`` (preprocess.py:328-530) for an injected
    ``numpy.random.Generator`` (the reference's own module-level generator is unseeded).

    ``interaction_dict``: ``{user: [items...]}`` in the reference's key and list order; ``kg``: either the reference-ordered
    ``{head: [(relation, tail), ...]}`` dict or a ``CKG`` (then the order is derived, ``kg_dict_reference_order``)."""

    def __init__(self, interaction_dict: dict, kg, item_num: int, node_num: int, rng: np.random.Generator, cf_batch_size: int = 256,
                 kg_batch_size: int = 512):
        self.rng = rng
        self.item_num, self.node_num = int(item_num), int(node_num)
        self.cf_batch_size, self.kg_batch_size = int(cf_batch_size), int(kg_batch_size)
        self._users = list(interaction_dict.keys())
        self._items = {u: list(v) for u, v in interaction_dict.items()}
        self._item_sets = {u: set(v) for u, v in self._items.items()}
        if isinstance(kg, CKG):
            key_heads, ptr, rels, tails = kg_dict_reference_order(kg)
            rt = np.stack([rels, tails], axis=1).tolist()
            kg = {int(h): [tuple(x) for x in rt[ptr[i] : ptr[i + 1]]] for i, h in enumerate(key_heads)}
        self._heads = list(kg.keys())
        self._triples = {h: list(v) for h, v in kg.items()}
        self._triple_sets = {h: set(v) for h, v in self._triples.items()}

    def generate_cf_batch(self):
        """-> (users, positive items, negative items), int64 arrays of length ``cf_batch_size`` (preprocess.py:380-415)."""
        rng, b = self.rng, self.cf_batch_size
        users = rng.choice(self._users, size=b, replace=b > len(self._users))
        pos, neg = [], []
        for u in users.tolist():
            items = self._items[u]
            pos.append(items[rng.integers(low=0, high=len(items), size=1)[0]])  # one draw: the set of size 1 fills at once
            seen = self._item_sets[u]
            while True:
                cand = int(rng.integers(low=0, high=self.item_num, size=1)[0])
                if cand not in seen:
                    neg.append(cand)
                    break
        return users.astype(np.int64), np.asarray(pos, np.int64), np.asarray(neg, np.int64)

    def generate_kg_batch(self):
        """-> (heads, relations, positive tails, negative tails) (preprocess.py:484-530)."""
        rng, b = self.rng, self.kg_batch_size
        heads = rng.choice(a=self._heads, size=b, replace=b > len(self._heads)).tolist()
        rels, pos, neg = [], [], []
        for h in heads:
            trip = self._triples[h]
            r, t = trip[rng.integers(low=0, high=len(trip))]
            rels.append(r)
            pos.append(t)
            seen = self._triple_sets[h]
            while True:
                cand = int(rng.integers(low=0, high=self.node_num, size=1)[0])
                if (r, cand) not in seen:
                    neg.append(cand)
                    break
        return np.asarray(heads, np.int64), np.asarray(rels, np.int64), np.asarray(pos, np.int64), np.asarray(neg, np.int64)


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
