"""Collaborative-knowledge-graph (CKG) assembly on the host.

Produces, from raw (user, item) interactions and (head, relation, tail) KG triples, exactly the
arrays the reference's ``Preprocess.run`` exposes to the KGAT model API:

* ``adjacency_relations``  (reference ``preprocess.py:177-222``)
* the head/tail-sorted edge list ``all_heads / all_relation_indices / all_tails / all_values``
  (reference ``preprocess.py:268-326``)
* the initial attentive matrix = sum of the per-relation "bi-normalised Laplacians"
  (reference ``preprocess.py:224-246, 628-634``)

The reference builds these with Python loops over every edge; here the same result is produced with
vectorised numpy so that Amazon-book-sized graphs (6.3 M edges) assemble in seconds.  Conventions
that must be preserved for parity:

* users occupy node ids ``[0, U)``; entity ``e`` is node ``U + e``; items are entities ``[0, I)``.
* with ``R0`` KG relation types, adjacency relation ids are: interaction ``0``, inverse interaction
  ``R0 + 1``, KG relation ``k`` -> ``k + 1``, its inverse -> ``k + 2 + R0``.
* each "Laplacian" is the *transpose* of its adjacency matrix scaled by the adjacency row degree,
  ``L = (D^-1/2 A)^T D^-1/2`` so ``L[j, i] = A[i, j] / deg_A(i)`` (computed as the float64 product
  ``deg^-0.5 * deg^-0.5`` and rounded to float32 exactly like the reference).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ---------------------------------------------------------------------------------------------
# Optional device assist: the sort / unique steps of the assembly dominate
# at 10^7-10^8 edges (numpy: minutes).  They are pure functions of their input, so running them with
# torch on a CUDA device gives bit-identical arrays (tested); everything that involves floating point
# (the 1/deg values) is still computed exactly as the numpy path does.  ``ACCEL`` = "auto" uses the GPU
# for arrays above ``ACCEL_MIN`` elements when one is visible, "never" / "always" force a path.
# ---------------------------------------------------------------------------------------------
ACCEL = "auto"
ACCEL_MIN = 2_000_000


def _use_gpu(n: int) -> bool:
    if ACCEL == "never":
        return False
    try:
        import torch

        if not torch.cuda.is_available():
            return False
    except Exception:  # pragma: no cover
        return False
    return ACCEL == "always" or n >= ACCEL_MIN


def stable_argsort(keys: np.ndarray) -> np.ndarray:
    if _use_gpu(keys.shape[0]):
        import torch

        return torch.sort(torch.from_numpy(np.ascontiguousarray(keys)).cuda(), stable=True).indices.cpu().numpy()
    return np.argsort(keys, kind="stable")


def sorted_unique(keys: np.ndarray, return_index: bool = False):
    """np.unique(keys[, return_index=True]) (first occurrence index of every unique key, keys need not be sorted)."""
    if _use_gpu(keys.shape[0]):
        import torch

        k = torch.from_numpy(np.ascontiguousarray(keys)).cuda()
        ks, order = torch.sort(k, stable=True)
        first = torch.ones(ks.shape[0], dtype=torch.bool, device=ks.device)
        first[1:] = ks[1:] != ks[:-1]
        uniq = ks[first].cpu().numpy()
        if return_index:
            return uniq, order[first].cpu().numpy()
        return uniq
    return np.unique(keys, return_index=True) if return_index else np.unique(keys)


def lexsort2(minor: np.ndarray, major: np.ndarray) -> np.ndarray:
    """np.lexsort((minor, major)): stable order by ``major`` then ``minor``."""
    if _use_gpu(major.shape[0]):
        import torch

        mi = torch.from_numpy(np.ascontiguousarray(minor)).cuda()
        ma = torch.from_numpy(np.ascontiguousarray(major)).cuda()
        o1 = torch.sort(mi, stable=True).indices
        o2 = torch.sort(ma[o1], stable=True).indices
        return o1[o2].cpu().numpy()
    return np.lexsort((minor, major))


@dataclass
class CKG:
    """Host-side CKG in the reference's conventions (all arrays numpy)."""

    user_num: int
    entity_num: int
    item_num: int
    kg_relation_num: int  # R0: KG relation types (without the interaction relation)
    adjacency_relations: list[int]  # 2*R0 + 2 relation ids, reference order
    # head/tail-sorted edge list (preprocess.py:268-326)
    heads: np.ndarray  # int32 (nnz,)
    relations: np.ndarray  # int64 (nnz,)
    tails: np.ndarray  # int32 (nnz,)
    values: np.ndarray  # float32 (nnz,)
    # initial attentive matrix, (row, col)-sorted, duplicates merged (preprocess.py:628-634)
    att_rows: np.ndarray  # int64
    att_cols: np.ndarray  # int64
    att_vals: np.ndarray  # float32
    # interaction splits: per-user item lists (train / validation / test)
    train_interactions: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int64))
    train_dict: dict[int, list[int]] = field(default_factory=dict)
    validation_dict: dict[int, list[int]] = field(default_factory=dict)
    test_dict: dict[int, list[int]] = field(default_factory=dict)

    @property
    def node_num(self) -> int:
        return self.user_num + self.entity_num

    @property
    def relation_num(self) -> int:
        return len(self.adjacency_relations)

    @property
    def nnz(self) -> int:
        return int(self.heads.shape[0])


def _laplacian_coo(rows: np.ndarray, cols: np.ndarray, node_num: int):
    """One relation's 'bi-normalised Laplacian' (preprocess.py:234-244) as sorted, merged COO.

    Input: adjacency COO (rows, cols) with unit values (duplicates add up, as in scipy).
    Output: (l_rows, l_cols, l_vals float64) sorted by (row, col) with duplicates summed.
    """
    if rows.size == 0:
        z = np.zeros(0, np.int64)
        return z, z, np.zeros(0, np.float64)
    deg = np.bincount(rows, minlength=node_num).astype(np.float64)
    with np.errstate(divide="ignore"):
        s = np.power(deg, -0.5)
    s[np.isinf(s)] = 0.0
    # (D^-1/2 A): value s[row]; transpose; times D^-1/2 on the right: value * s[row] again.
    vals = (1.0 * s[rows]) * s[rows]
    l_rows, l_cols = cols.astype(np.int64), rows.astype(np.int64)
    key = l_rows * node_num + l_cols
    order = stable_argsort(key)
    key, vals = key[order], vals[order]
    uniq, start = sorted_unique(key, return_index=True)
    if uniq.size != key.size:  # duplicate adjacency entries: scipy sums them
        vals = np.add.reduceat(vals, start)
    return uniq // node_num, uniq % node_num, vals


def build_ckg(
    user_num: int,
    entity_num: int,
    item_num: int,
    kg_relation_num: int,
    interactions: np.ndarray,
    triples: np.ndarray,
) -> CKG:
    """Assemble the CKG arrays from interactions ``(M, 2) [user, item]`` and triples ``(T, 3)
    [head_entity, kg_relation, tail_entity]`` (entity-local ids).  Restates ``Preprocess``
    steps P1-P4 (preprocess.py:157-326, 628-634)."""
    interactions = np.asarray(interactions, dtype=np.int64).reshape(-1, 2)
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    n = user_num + entity_num
    r0 = kg_relation_num

    adj: list[tuple[np.ndarray, np.ndarray]] = []
    adjacency_relations: list[int] = []
    u = interactions[:, 0]
    p = interactions[:, 1] + user_num
    adj.append((u, p))
    adjacency_relations.append(0)
    adj.append((p, u))
    adjacency_relations.append(r0 + 1)
    for k in range(r0):
        sel = triples[:, 1] == k
        h = triples[sel, 0] + user_num
        t = triples[sel, 2] + user_num
        adj.append((h, t))
        adjacency_relations.append(k + 1)
        adj.append((t, h))
        adjacency_relations.append(k + 2 + r0)

    lap = [_laplacian_coo(r, c, n) for r, c in adj]

    # ---- edge list: concatenated Laplacians, grouped by head, tail-sorted within head ----------
    heads = np.concatenate([l[0] for l in lap])
    tails = np.concatenate([l[1] for l in lap])
    vals = np.concatenate([l[2] for l in lap])
    rels = np.concatenate(
        [np.full(l[0].shape[0], rid, dtype=np.int64) for l, rid in zip(lap, adjacency_relations)]
    )
    order = stable_argsort(heads * n + tails)  # == np.lexsort((tails, heads)); stable: ties keep Laplacian order
    heads, tails, vals, rels = heads[order], tails[order], vals[order], rels[order]

    # ---- initial attentive matrix: sum of Laplacians in float64, then float32 ------------------
    key = heads * n + tails  # already (row, col)-sorted
    uniq, start = sorted_unique(key, return_index=True)
    if uniq.size != key.size:
        att_vals64 = np.add.reduceat(vals, start)
    else:
        att_vals64 = vals
    att_rows, att_cols = uniq // n, uniq % n

    return CKG(
        user_num=user_num,
        entity_num=entity_num,
        item_num=item_num,
        kg_relation_num=r0,
        adjacency_relations=adjacency_relations,
        heads=heads.astype(np.int32),
        relations=rels,
        tails=tails.astype(np.int32),
        values=vals.astype(np.float32),
        att_rows=att_rows,
        att_cols=att_cols,
        att_vals=att_vals64.astype(np.float32),
    )


def interaction_dict(pairs: np.ndarray, user_num: int | None = None) -> dict[int, list[int]]:
    """``{user: [items...]}`` in first-seen order (preprocess.py:131-134 builds per-user lists)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    out: dict[int, list[int]] = {} if user_num is None else {int(u): [] for u in range(user_num)}
    if pairs.size:
        order = stable_argsort(pairs[:, 0])
        sp = pairs[order]
        users, start = np.unique(sp[:, 0], return_index=True)
        bounds = list(start) + [sp.shape[0]]
        for i, usr in enumerate(users):
            out[int(usr)] = sp[bounds[i] : bounds[i + 1], 1].tolist()
    return out
