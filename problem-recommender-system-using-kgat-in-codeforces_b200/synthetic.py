"""Seeded synthetic CKGs of the BASELINE.json shapes.

There is no network and the reference's two large data blobs are absent, so every configuration is
measured on synthetic graphs that follow the reference's node/relation conventions (see ``ckg.py``).
Everything is driven by ``numpy.random.default_rng(seed)`` with ``seed = 2024`` (the reference's
``SEED``, ``src/constants.py:4``).

Degree laws: item popularity follows a shifted Zipf ``p(rank) ~ (rank + shift)^-a`` (a = 1.3),
users are uniform; KG tails follow a shifted Zipf over the non-item entities, rotated per relation
so every relation has its own hub entities.  Interactions are unique ``(user, item)`` pairs and are
split per user 72 / 8 / 20 into train / validation / test (reference: ``preprocess.py:77-90``
0.8 x 0.9 / 0.8 x 0.1 / 0.2); the CKG is built from the training interactions only, as in
``Preprocess.run("training")``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .ckg import CKG, build_ckg, interaction_dict, lexsort2, sorted_unique

SEED = 2024


@dataclass(frozen=True)
class Shape:
    name: str
    user_num: int
    item_num: int
    entity_num: int
    kg_relation_num: int
    triple_num: int
    interaction_num: int
    schema: str = "generic"  # or "codeforces"


# BASELINE.json configs[0..4]
SHAPES: dict[str, Shape] = {
    # C1 Codeforces small: 500 users, ~9.5k problems, + contests/divisions/tags/ratings
    "codeforces-sm": Shape("codeforces-sm", 500, 9_500, 9_500 + 1_913 + 5 + 37 + 28, 4, 0, 500 * 1_000, "codeforces"),
    # C2 full Codeforces
    "codeforces-full": Shape(
        "codeforces-full", 200_000, 10_000, 10_000 + 1_913 + 5 + 37 + 28, 4, 0, 20_000_000, "codeforces"
    ),
    # C3 Amazon-book
    "amazon-book": Shape("amazon-book", 70_679, 24_915, 88_572, 39, 2_557_746, 847_733),
    # C4 Yelp2018
    "yelp2018": Shape("yelp2018", 45_919, 45_538, 90_961, 42, 1_853_704, 1_185_068),
    # C5 scaled (edge count given directly: 200 M directed CKG edges)
    "scaled": Shape("scaled", 10_000_000, 1_000_000, 1_000_000 + 16, 31, 10_000_000, 125_000_000),
    # tiny shapes for tests
    "tiny": Shape("tiny", 40, 60, 90, 3, 150, 400),
    "small": Shape("small", 300, 500, 800, 5, 2_500, 6_000),
}


def _zipf_probs(n: int, a: float, shift: float) -> np.ndarray:
    p = np.power(np.arange(1, n + 1, dtype=np.float64) + shift, -a)
    return p / p.sum()


def _unique_pairs(rng, n_target: int, draw, key_mod: int) -> np.ndarray:
    """Draw until ``n_target`` unique (a, b) pairs exist; returns (n_target, 2) int64."""
    keys = np.zeros(0, np.int64)
    want = n_target
    for _ in range(64):
        a, b = draw(int(want * 1.3) + 16)
        keys = sorted_unique(np.concatenate([keys, a.astype(np.int64) * key_mod + b.astype(np.int64)]))
        if keys.size >= n_target:
            break
        want = n_target - keys.size
    if keys.size > n_target:
        keys = np.sort(rng.choice(keys, size=n_target, replace=False))
    return np.stack([keys // key_mod, keys % key_mod], axis=1)


def _interactions(rng, shape: Shape, zipf_a: float, zipf_shift: float) -> np.ndarray:
    n_target = min(shape.interaction_num, shape.user_num * shape.item_num // 2)
    p_item = _zipf_probs(shape.item_num, zipf_a, zipf_shift)
    item_perm = rng.permutation(shape.item_num)
    cdf = np.cumsum(p_item)
    cdf[-1] = 1.0

    def draw(k):
        users = rng.integers(0, shape.user_num, size=k)
        items = item_perm[np.searchsorted(cdf, rng.random(k), side="right").clip(max=shape.item_num - 1)]
        return users, items

    pairs = _unique_pairs(rng, n_target, draw, shape.item_num)
    # every user needs at least one interaction (reference samplers index interaction_dict[user])
    missing = np.setdiff1d(np.arange(shape.user_num), pairs[:, 0])
    if missing.size:
        extra = np.stack([missing, item_perm[rng.integers(0, min(64, shape.item_num), size=missing.size)]], axis=1)
        pairs = np.concatenate([pairs, extra])
        pairs = pairs[lexsort2(pairs[:, 1], pairs[:, 0])]
    return pairs


def _generic_triples(rng, shape: Shape, zipf_a: float) -> np.ndarray:
    n_other = shape.entity_num - shape.item_num
    r = shape.kg_relation_num
    p_rel = _zipf_probs(r, 0.8, 2.0)
    p_tail = _zipf_probs(n_other, zipf_a, 5.0)
    cdf_tail = np.cumsum(p_tail)
    cdf_tail[-1] = 1.0
    stride = max(1, n_other // max(r, 1))

    def draw(k):
        rel = rng.choice(r, size=k, p=p_rel)
        head = rng.integers(0, shape.item_num, size=k)
        rank = np.searchsorted(cdf_tail, rng.random(k), side="right").clip(max=n_other - 1)
        tail = shape.item_num + (rank + rel * stride) % n_other
        return head * r + rel, tail  # pack (head, rel) as the first key

    pairs = _unique_pairs(rng, shape.triple_num, draw, shape.entity_num)
    return np.stack([pairs[:, 0] // r, pairs[:, 0] % r, pairs[:, 1]], axis=1)


def _codeforces_triples(rng, shape: Shape) -> np.ndarray:
    """Codeforces schema (reference ``kg_triplets_generator.py:136-197``, ``type.py:90-94``):
    entities = problems, then contests, divisions, tags, ratings; relations TAGGED=0,
    HAS_DIFFICULTY=1, IN_CONTEST=2, HAS_CONTEST_DIVISION=3."""
    n_items = shape.item_num
    n_contests, n_div, n_tags, n_ratings = 1_913, 5, 37, 28
    n_contests = min(n_contests, shape.entity_num - n_items - n_div - n_tags - n_ratings)
    c0 = n_items
    d0 = c0 + n_contests
    t0 = d0 + n_div
    g0 = t0 + n_tags
    problems = np.arange(n_items)
    contest_of = c0 + np.sort(rng.integers(0, n_contests, size=n_items))
    in_contest = np.stack([problems, np.full(n_items, 2), contest_of], axis=1)
    used_contests = np.unique(contest_of)
    division = np.stack(
        [used_contests, np.full(used_contests.size, 3), d0 + rng.integers(0, n_div, size=used_contests.size)], axis=1
    )
    n_tag_per = rng.poisson(2.6, size=n_items).clip(0, 8)
    tag_heads = np.repeat(problems, n_tag_per)
    p_tag = _zipf_probs(n_tags, 1.0, 2.0)
    tag_tails = t0 + rng.choice(n_tags, size=tag_heads.size, p=p_tag)
    tagged = np.unique(np.stack([tag_heads, np.zeros_like(tag_heads), tag_tails], axis=1), axis=0)
    has_rating = rng.random(n_items) < 0.9
    rated = np.stack(
        [problems[has_rating], np.full(int(has_rating.sum()), 1), g0 + rng.integers(0, n_ratings, size=int(has_rating.sum()))],
        axis=1,
    )
    return np.concatenate([in_contest, division, tagged, rated]).astype(np.int64)


def split_interactions(rng, pairs: np.ndarray, user_num: int):
    """Per-user 72 / 8 / 20 split (at least one training item per user)."""
    jitter = rng.random(pairs.shape[0])
    order = lexsort2(jitter, pairs[:, 0])
    sp = pairs[order]
    counts = np.bincount(sp[:, 0], minlength=user_num)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    pos = np.arange(sp.shape[0]) - np.repeat(starts, counts)
    n_u = np.repeat(counts, counts)
    n_train = np.maximum(1, np.floor(0.72 * n_u + 0.5).astype(np.int64))
    n_val = np.floor(0.08 * n_u + 0.5).astype(np.int64)
    is_train = pos < n_train
    is_val = (~is_train) & (pos < n_train + n_val)
    is_test = ~(is_train | is_val)
    return sp[is_train], sp[is_val], sp[is_test]


def make_ckg(
    shape: str | Shape,
    seed: int = SEED,
    zipf_a: float = 1.3,
    zipf_shift: float = 50.0,
    duplicate_pairs: int = 0,
    with_dicts: bool = True,
) -> CKG:
    """Generate a synthetic CKG.  ``duplicate_pairs`` > 0 adds that many extra triples whose
    (head, tail) already exists under another relation (the duplicate-(h, t) stress case: the
    reference sums such entries before the row softmax, ``model.py:364``)."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    rng = np.random.default_rng(seed)
    pairs = _interactions(rng, shape, zipf_a, zipf_shift)
    if shape.schema == "codeforces":
        triples = _codeforces_triples(rng, shape)
    else:
        triples = _generic_triples(rng, shape, zipf_a)
    if duplicate_pairs > 0 and triples.shape[0] > 0 and shape.kg_relation_num > 1:
        pick = triples[rng.integers(0, triples.shape[0], size=duplicate_pairs)].copy()
        pick[:, 1] = (pick[:, 1] + 1 + rng.integers(0, shape.kg_relation_num - 1, size=duplicate_pairs)) % shape.kg_relation_num
        triples = np.unique(np.concatenate([triples, pick]), axis=0)
    train, val, test = split_interactions(rng, pairs, shape.user_num)
    g = build_ckg(shape.user_num, shape.entity_num, shape.item_num, shape.kg_relation_num, train, triples)
    g.train_interactions = train
    if with_dicts:
        g.train_dict = interaction_dict(train, shape.user_num)
        g.validation_dict = interaction_dict(val, shape.user_num)
        g.test_dict = interaction_dict(test, shape.user_num)
    return g


def make_edges_only(node_num: int, nnz: int, relation_num: int, seed: int = SEED, zipf_a: float = 1.1):
    """C5-style graph given directly by its edge count: returns (heads, rels, tails) sorted by
    (head, tail) with unique (h, t), symmetric relation ids, for per-step propagation benchmarks."""
    rng = np.random.default_rng(seed)
    half = nnz // 2
    p = _zipf_probs(node_num, zipf_a, 1000.0)
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    perm_mul = 2654435761 % node_num | 1

    def draw(k):
        a = rng.integers(0, node_num, size=k)
        b = (np.searchsorted(cdf, rng.random(k), side="right").clip(max=node_num - 1) * perm_mul) % node_num
        return a, b

    pr = _unique_pairs(rng, half, draw, node_num)
    pr = pr[pr[:, 0] != pr[:, 1]]
    rel = rng.integers(0, relation_num // 2, size=pr.shape[0])
    h = np.concatenate([pr[:, 0], pr[:, 1]])
    t = np.concatenate([pr[:, 1], pr[:, 0]])
    r = np.concatenate([rel, rel + relation_num // 2])
    key = h * node_num + t
    _, first = sorted_unique(key, return_index=True)
    h, t, r = h[first], t[first], r[first]
    return h.astype(np.int32), r.astype(np.int64), t.astype(np.int32)
