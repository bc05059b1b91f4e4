"""CPU oracle for the KGAT hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU (fp32) / numpy restatement of the reference's algorithm for the path named by
BASELINE.json's ``north_star``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package never does
(the product path raises when the CUDA library is missing).

Parity status: **pinned against the reference itself.**  The reference has no tests or golden
vectors of its own, so ``oracle/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` in the build container, runs them on seeded inputs and
stores inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors (CPU, no GPU needed).

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
State is passed as a dict keyed by the reference's ``state_dict`` names so reference checkpoints
plug in directly.  The quirks Q1-Q5 of SURVEY.md section 0 are reproduced on purpose.
"""

from __future__ import annotations

import math
from collections import OrderedDict, defaultdict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------

MHA_HEADS = 8  # src/model/KGAT/multi_head_attention.py:6 (head_num=8, dropout=0.1 hard-coded)
MHA_DROPOUT = 0.1
LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, src/model/KGAT/aggregator.py:23
NORM_EPS = 1e-12  # F.normalize default eps, aggregator.py:65
LN_EPS = 1e-5  # nn.LayerNorm default eps, multi_head_attention.py:17


def init_params(node_num: int, relation_num: int, dim: int = 64, layer_size=(64, 32, 16), seed: int = 2024):
    """Random fp32 parameters with the reference's ``state_dict`` key set and init distributions
    (src/model/KGAT/model.py:59-105, aggregator.py:25-35, multi_head_attention.py:13-29).
    Not RNG-stream identical to constructing the reference module; parity tests load reference
    state_dicts from the golden files instead."""
    g = torch.Generator().manual_seed(seed)

    def xavier(*shape):
        if len(shape) == 2:
            fan_out, fan_in = shape
        else:  # (R, d, d): torch fan computation for >2 dims
            rf = math.prod(shape[2:])
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def bias(fan_in, n):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(n, generator=g) * 2 - 1) * b

    p = OrderedDict()
    p["_trans_matrix"] = xavier(relation_num, dim, dim)
    p["_user_entity_embedding.weight"] = xavier(node_num, dim)
    p["_relation_embedding.weight"] = xavier(relation_num, dim)
    dims = [dim, *layer_size]
    for l in range(len(layer_size)):
        for k in (1, 2):
            p[f"_aggregator_layers.{l}.linear{k}.weight"] = xavier(dims[l + 1], dims[l])
            p[f"_aggregator_layers.{l}.linear{k}.bias"] = bias(dims[l], dims[l + 1])
    for name in ("_query_weight", "_key_weight", "_value_weight", "_output"):
        p[f"_multi_head_attention.{name}.weight"] = xavier(dim, dim)
        p[f"_multi_head_attention.{name}.bias"] = bias(dim, dim)
    p["_multi_head_attention._layer_norm.weight"] = torch.ones(dim)
    p["_multi_head_attention._layer_norm.bias"] = torch.zeros(dim)
    return p


def layer_count(params) -> int:
    n = 0
    while f"_aggregator_layers.{n}.linear1.weight" in params:
        n += 1
    return n


# ----------------------------------------------------------------------------------------------
# A3 / A4: attentive propagation
# ----------------------------------------------------------------------------------------------


def sparse_coo(rows, cols, vals, n: int) -> torch.Tensor:
    """COO tensor as the reference builds it (preprocess.py:630-634, model.py:359-363)."""
    idx = torch.stack([torch.as_tensor(rows, dtype=torch.long), torch.as_tensor(cols, dtype=torch.long)])
    return torch.sparse_coo_tensor(idx, torch.as_tensor(vals, dtype=torch.float32), size=(n, n))


def aggregator_forward(ego, att, w1, b1, w2, b2, drop_mask=None):
    """One bi-interaction layer, src/model/KGAT/aggregator.py:37-65.

    ``drop_mask`` (same shape as the output, entries 0 or 1/(1-p)) stands in for
    ``nn.Dropout`` (aggregator.py:62); ``None`` = eval mode.  Returns (out, side)."""
    side = torch.matmul(att, ego)  # :54
    s = F.leaky_relu(F.linear(ego + side, w1, b1), LEAKY_SLOPE)  # :57
    m = F.leaky_relu(F.linear(ego * side, w2, b2), LEAKY_SLOPE)  # :58
    out = s + m  # :59
    if drop_mask is not None:
        out = out * drop_mask  # :62
    return F.normalize(out, p=2.0, dim=1, eps=NORM_EPS), side  # :65


def propagate(params, att, drop_masks=None):
    """src/model/KGAT/model.py:124-140: returns the list [E0, E1, ..., EL] (cat = all embeddings)."""
    ego = params["_user_entity_embedding.weight"]
    outs = [ego]
    for l in range(layer_count(params)):
        ego, _ = aggregator_forward(
            ego,
            att,
            params[f"_aggregator_layers.{l}.linear1.weight"],
            params[f"_aggregator_layers.{l}.linear1.bias"],
            params[f"_aggregator_layers.{l}.linear2.weight"],
            params[f"_aggregator_layers.{l}.linear2.bias"],
            None if drop_masks is None else drop_masks[l],
        )
        outs.append(ego)
    return outs


def all_embeddings(params, att, drop_masks=None):
    return torch.cat(propagate(params, att, drop_masks), dim=1)  # model.py:140


# ----------------------------------------------------------------------------------------------
# A5: BPR loss, A6: TransR loss
# ----------------------------------------------------------------------------------------------


def l2_mean(x):
    """src/model/KGAT/model.py:142-163."""
    return torch.mean(torch.sum(torch.pow(x, 2), dim=1) / 2.0)


def bpr_loss_from_table(table, users, pos, neg, reg: float = 1e-5):
    """src/model/KGAT/model.py:189-202 (ids used as given: no user_num offset, quirk Q4)."""
    u = table[users.long()]
    p = table[pos.long()]
    n = table[neg.long()]
    ps = torch.sum(u * p, dim=1)
    ns = torch.sum(u * n, dim=1)
    return -F.logsigmoid(ps - ns).mean() + reg * (l2_mean(u) + l2_mean(p) + l2_mean(n))


def cf_loss(params, att, users, pos, neg, reg: float = 1e-5, drop_masks=None):
    """src/model/KGAT/model.py:165-202."""
    return bpr_loss_from_table(all_embeddings(params, att, drop_masks), users, pos, neg, reg)


def kg_loss(params, heads, rels, pos_tails, neg_tails, reg: float = 1e-5):
    """TransR loss, src/model/KGAT/model.py:204-261 (row-vector convention e @ W_r)."""
    emb = params["_user_entity_embedding.weight"]
    e_r = params["_relation_embedding.weight"][rels.long()]
    w = params["_trans_matrix"][rels.long()]  # (B, d, k) materialised, as the reference does
    h = torch.matmul(emb[heads.long()].unsqueeze(1), w).squeeze(1)
    p = torch.matmul(emb[pos_tails.long()].unsqueeze(1), w).squeeze(1)
    n = torch.matmul(emb[neg_tails.long()].unsqueeze(1), w).squeeze(1)
    ps = torch.sum(torch.pow(h + e_r - p, 2), dim=1)
    ns = torch.sum(torch.pow(h + e_r - n, 2), dim=1)
    loss = -F.logsigmoid(ns - ps).mean()
    return loss + reg * (l2_mean(h) + l2_mean(e_r) + l2_mean(p) + l2_mean(n))


# ----------------------------------------------------------------------------------------------
# A7 / A8: attention refresh
# ----------------------------------------------------------------------------------------------


def mha_forward(params, x_head, e_rel, x_tail, head_mask=None):
    """src/model/KGAT/multi_head_attention.py:35-58.  ``head_mask`` (B, 8) with entries 0 or
    1/(1-0.1) replaces ``nn.Dropout`` on the (B, 8, 1, 1) attention tensor; ``None`` = eval."""
    pre = "_multi_head_attention."
    b = x_head.size(0)
    d = x_head.size(1)
    depth = d // MHA_HEADS

    def split(x):
        return x.view(b, -1, MHA_HEADS, depth).transpose(1, 2)

    q = split(F.linear(x_head, params[pre + "_query_weight.weight"], params[pre + "_query_weight.bias"]))
    k = split(
        F.linear(e_rel.unsqueeze(0).expand(b, -1), params[pre + "_key_weight.weight"], params[pre + "_key_weight.bias"])
    )
    v = split(F.linear(x_tail, params[pre + "_value_weight.weight"], params[pre + "_value_weight.bias"]))
    att = torch.matmul(q, k.transpose(-2, -1)) / (depth**0.5)  # (B, 8, 1, 1)
    att = torch.softmax(att, dim=-1)  # softmax over a size-1 axis == 1.0 (quirk Q1)
    if head_mask is not None:
        att = att * head_mask.view(b, MHA_HEADS, 1, 1)
    o = torch.matmul(att, v).transpose(1, 2).contiguous().view(b, -1, d)
    o = F.linear(o, params[pre + "_output.weight"], params[pre + "_output.bias"])
    return F.layer_norm(o, (d,), params[pre + "_layer_norm.weight"], params[pre + "_layer_norm.bias"], LN_EPS)


def attention_by_relation(params, heads, tails, rel: int, node_num: int, head_mask=None):
    """src/model/KGAT/model.py:263-316 for the edges of one relation."""
    emb = params["_user_entity_embedding.weight"]
    e_r = params["_relation_embedding.weight"][rel]
    w = params["_trans_matrix"][rel]
    x_h = torch.matmul(emb[heads], w)
    x_t = torch.matmul(emb[tails], w)
    out = mha_forward(params, x_h, e_r, x_t, head_mask)
    score = torch.sum(torch.tanh(out.squeeze(1)), dim=1)
    deg_h = torch.bincount(heads, minlength=node_num)
    deg_t = torch.bincount(tails, minlength=node_num)
    weight = 1.0 / (torch.log1p(deg_h[heads]) + torch.log1p(deg_t[tails]))
    return score * weight


def attention_refresh(params, heads, rels, tails, relation_indices, node_num: int, head_masks=None):
    """src/model/KGAT/model.py:318-366.  Returns the new attentive matrix as a *coalesced* COO
    (rows, cols int64; vals fp32): duplicates (h, t) are summed before the row softmax and the
    result is (row, col)-sorted (``torch.sparse.softmax`` on CPU, quirk Q3).
    ``head_masks``: optional dict {relation id: (n_r, 8) mask} (quirk Q2: dropout is live when the
    model is in train mode)."""
    heads = torch.as_tensor(heads).long()
    tails = torch.as_tensor(tails).long()
    rels = torch.as_tensor(rels).long()
    rows, cols, vals = [], [], []
    for r in torch.as_tensor(relation_indices).long().tolist():
        sel = torch.where(rels == r)[0]
        h, t = heads[sel], tails[sel]
        hm = None if head_masks is None else head_masks.get(r)
        rows.append(h)
        cols.append(t)
        vals.append(attention_by_relation(params, h, t, r, node_num, hm))
    m = torch.sparse_coo_tensor(torch.stack([torch.cat(rows), torch.cat(cols)]), torch.cat(vals), size=(node_num, node_num))
    m = torch.sparse.softmax(m, dim=1)
    return m.indices()[0].clone(), m.indices()[1].clone(), m.values().clone()


def segment_softmax_coalesced(rows, cols, vals, node_num: int):
    """numpy restatement of coalesce + row softmax (what model.py:364 does), used to cross-check
    ``torch.sparse.softmax`` and as the CSR-order oracle for the segmented-softmax kernel."""
    rows = np.asarray(rows, np.int64)
    cols = np.asarray(cols, np.int64)
    vals = np.asarray(vals, np.float32)
    key = rows * node_num + cols
    order = np.argsort(key, kind="stable")
    key, v = key[order], vals[order]
    uk, start = np.unique(key, return_index=True)
    merged = np.add.reduceat(v.astype(np.float32), start).astype(np.float32) if v.size else v
    r = uk // node_num
    out = np.empty_like(merged)
    row_ids, row_start = np.unique(r, return_index=True)
    bounds = list(row_start) + [r.size]
    for i in range(len(row_ids)):
        seg = merged[bounds[i] : bounds[i + 1]]
        e = np.exp(seg - seg.max())
        out[bounds[i] : bounds[i + 1]] = e / e.sum(dtype=np.float32)
    return r, uk % node_num, out


# ----------------------------------------------------------------------------------------------
# A9: predict, M1: ranking metrics
# ----------------------------------------------------------------------------------------------


def predict_scores(params, att, users, items):
    """src/model/KGAT/model.py:368-391 (eval mode: no dropout)."""
    table = all_embeddings(params, att)
    return torch.matmul(table[users.long()], table[items.long()].transpose(0, 1))


def rank_items(scores: torch.Tensor, train_dict, user_ids):
    """src/utils/metrics_calculator.py:114-121: mask the training positives to -inf (raw item ids
    index columns), full descending sort.  CPU ``torch.sort`` resolves exact ties lowest index
    first (stable).  Returns the int64 rank-index matrix."""
    s = scores.clone()
    for i, u in enumerate(user_ids):
        items = train_dict[int(u)]
        if len(items):
            s[i][items] = -np.inf
    _, idx = torch.sort(s, descending=True, stable=True)
    return idx


def metrics_at_k(scores, train_dict, test_dict, user_ids, n_items: int, k_list):
    """src/utils/metrics_calculator.py:84-131 (precision / recall / nDCG at K per user)."""
    idx = rank_items(scores, train_dict, user_ids).numpy()
    pos = np.zeros((len(user_ids), n_items), np.float32)
    for i, u in enumerate(user_ids):
        pos[i][test_dict[int(u)]] = 1
    hits = np.take_along_axis(pos, idx, axis=1)
    out = {}
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in k_list:
            hk = hits[:, :k]
            disc = np.log2(np.arange(2, k + 2))
            dcg = np.sum((2**hk - 1) / disc, axis=1)
            ideal = np.flip(np.sort(hits), axis=1)[:, :k]
            idcg = np.sum((2**ideal - 1) / disc, axis=1)
            idcg[idcg == 0] = np.inf
            out[k] = {
                "precision": hk.mean(axis=1),
                "recall": hk.sum(axis=1) / hits.sum(axis=1),
                "ndcg": dcg / idcg,
            }
    return out


# ----------------------------------------------------------------------------------------------
# A10: Adam (torch.optim.Adam defaults, model.py:404-405)
# ----------------------------------------------------------------------------------------------


def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """One ``torch.optim.Adam`` update (single-tensor path of torch/optim/adam.py; no weight decay,
    no amsgrad).  Mutates and returns (param, exp_avg, exp_avg_sq); ``step`` is 1-based."""
    exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    step_size = lr / bc1
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-step_size)
    return param, exp_avg, exp_avg_sq


# ----------------------------------------------------------------------------------------------
# P1-P4: CKG construction (loop-level restatement; the product uses a vectorised builder)
# ----------------------------------------------------------------------------------------------


def build_ckg(user_num: int, entity_num: int, kg_relation_num: int, interactions, triples):
    """src/model/KGAT/preprocess.py:157-326 and 628-634 restated with scipy exactly as the
    reference composes it.  ``interactions`` (M, 2) [user, item]; ``triples`` (T, 3)
    [head, relation, tail] in entity-local ids.  Returns a dict with the reference's outputs."""
    import scipy.sparse as sp

    n = user_num + entity_num
    interactions = np.asarray(interactions, np.int64).reshape(-1, 2)
    triples = np.asarray(triples, np.int64).reshape(-1, 3)

    def pair(rows, cols):  # preprocess.py:157-175
        ones = [1.0] * len(rows)
        return (
            sp.coo_matrix((ones, (rows, cols)), shape=(n, n)),
            sp.coo_matrix((ones, (cols, rows)), shape=(n, n)),
        )

    mats, rel_ids = [], []
    a, a_inv = pair(interactions[:, 0], interactions[:, 1] + user_num)  # :194-202
    mats += [a, a_inv]
    rel_ids += [0, kg_relation_num + 1]
    for k in range(kg_relation_num):  # :207-219
        sel = triples[triples[:, 1] == k]
        a, a_inv = pair(sel[:, 0] + user_num, sel[:, 2] + user_num)
        mats += [a, a_inv]
        rel_ids += [k + 1, k + 2 + kg_relation_num]

    def bi_norm(m):  # :234-244
        row_sum = np.array(m.sum(axis=1))
        with np.errstate(divide="ignore"):
            d = np.power(row_sum, -0.5).flatten()
        d[np.isinf(d)] = 0.0
        dm = sp.diags(d)
        return dm.dot(m).transpose().dot(dm).tocoo()

    laps = [bi_norm(m) for m in mats]

    by_head = defaultdict(lambda: ([], [], []))  # :286-302
    for lid, lap in enumerate(laps):
        for r, c, v in zip(lap.row, lap.col, lap.data):
            e = by_head[int(r)]
            e[0].append(rel_ids[lid])
            e[1].append(int(c))
            e[2].append(v)
    heads, rels, tails, vals = [], [], [], []
    for h in sorted(by_head):  # :304-324
        r_l, t_l, v_l = by_head[h]
        order = np.argsort(t_l)
        heads += [h] * len(t_l)
        rels += np.array(r_l, np.int64)[order].tolist()
        tails += np.array(t_l, np.int64)[order].tolist()
        vals += np.array(v_l, np.float32)[order].tolist()

    total = sum(laps).tocoo()  # :629
    return {
        "adjacency_relations": rel_ids,
        "heads": np.array(heads, np.int32),
        "relations": np.array(rels, np.int64),
        "tails": np.array(tails, np.int32),
        "values": np.array(vals, np.float32),
        "att_rows": np.asarray(total.row, np.int64),
        "att_cols": np.asarray(total.col, np.int64),
        "att_vals": np.asarray(total.data).astype(np.float32),
    }


def kg_dict_from_edges(heads, rels, tails):
    """``kg_dict[head] = [(relation, tail), ...]`` (preprocess.py:248-266).  The reference fills it
    Laplacian by Laplacian; sampling draws only use the list as an indexable multiset, so for the
    RNG-stream parity test the golden fixture stores the reference's own list order."""
    d = defaultdict(list)
    for h, r, t in zip(np.asarray(heads).tolist(), np.asarray(rels).tolist(), np.asarray(tails).tolist()):
        d[h].append((r, t))
    return dict(d)


# ----------------------------------------------------------------------------------------------
# P5: samplers (sequential, RNG-stream compatible with the reference given the same Generator)
# ----------------------------------------------------------------------------------------------


def sample_cf_batch(rng, inter_dict, item_num: int, batch_size: int):
    """src/model/KGAT/preprocess.py:328-415 with an injected ``numpy.random.Generator`` (the
    reference's module-level ``rng`` is unseeded, quirk Q5)."""
    users_all = list(inter_dict.keys())
    users = rng.choice(users_all, size=batch_size, replace=batch_size > len(users_all))
    pos, neg = [], []
    for u in users:
        items = inter_dict[u]
        pos.append(items[rng.integers(low=0, high=len(items), size=1)[0]])
        while True:
            cand = rng.integers(low=0, high=item_num, size=1)[0]
            if cand not in items:
                neg.append(cand)
                break
    return np.asarray(users, np.int64), np.asarray(pos, np.int64), np.asarray(neg, np.int64)


def sample_kg_batch(rng, kg_dict, node_num: int, batch_size: int):
    """src/model/KGAT/preprocess.py:417-530."""
    heads_all = list(kg_dict.keys())
    heads = rng.choice(a=heads_all, size=batch_size, replace=batch_size > len(heads_all)).tolist()
    rel_b, pos_b, neg_b = [], [], []
    for h in heads:
        trip = kg_dict[h]
        i = rng.integers(low=0, high=len(trip))
        r, t = trip[i]
        rel_b.append(r)
        pos_b.append(t)
        while True:
            cand = rng.integers(low=0, high=node_num, size=1)[0]
            if (r, cand) not in trip:
                neg_b.append(cand)
                break
    return (
        np.asarray(heads, np.int64),
        np.asarray(rel_b, np.int64),
        np.asarray(pos_b, np.int64),
        np.asarray(neg_b, np.int64),
    )


# ----------------------------------------------------------------------------------------------
# canonical KGAT attention (north_star item (1)); NOT what the reference computes -- the oracle of model.score_mode = "kgat"
# ----------------------------------------------------------------------------------------------


def attention_refresh_kgat(params, heads, rels, tails, relation_indices, node_num: int):
    """The KGAT paper's pi(h, r, t) = (W_r e_t)^T tanh(W_r e_h + e_r) (arXiv 1905.07854, eq. 4) in the reference's row-vector
    convention x = e W_r (src/model/KGAT/model.py:291-298), then exactly the reference's assembly (model.py:355-366): per-relation
    batches concatenated into a COO, duplicates (h, t) summed by the coalesce inside ``torch.sparse.softmax``, row softmax.
    No degree weights, no multi-head attention.  Plain PyTorch; pinned by construction (it IS the formula), used to check the kernel."""
    heads = torch.as_tensor(heads).long()
    tails = torch.as_tensor(tails).long()
    rels = torch.as_tensor(rels).long()
    emb, rel_emb, w = params["_user_entity_embedding.weight"], params["_relation_embedding.weight"], params["_trans_matrix"]
    rows, cols, vals = [], [], []
    for r in torch.as_tensor(relation_indices).long().tolist():
        sel = torch.where(rels == r)[0]
        h, t = heads[sel], tails[sel]
        x_h = torch.matmul(emb[h], w[r])
        x_t = torch.matmul(emb[t], w[r])
        rows.append(h)
        cols.append(t)
        vals.append(torch.sum(x_t * torch.tanh(x_h + rel_emb[r]), dim=1))
    m = torch.sparse_coo_tensor(torch.stack([torch.cat(rows), torch.cat(cols)]), torch.cat(vals), size=(node_num, node_num))
    m = torch.sparse.softmax(m, dim=1)
    return m.indices()[0].clone(), m.indices()[1].clone(), m.values().clone()
