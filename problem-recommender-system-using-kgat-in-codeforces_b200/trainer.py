"""Epoch driver restating the reference's training-epoch body (src/model/KGAT/main.py:290-361):

    n_cf = len(train_interactions) // 256 + 1   CF steps  (forward, backward, Adam)      main.py:297-316
    n_kg = nnz // 512 + 1                       KG steps  (forward, backward, Adam)      main.py:324-345
    one attention refresh (model still in train() mode)                   main.py:350-361

Only the loop structure is restated (``main.py`` itself needs matplotlib and the crawled dataset);
every model call goes through the same public API the reference driver uses.  Batches are
pre-sampled (``sampler.BatchSampler``) so the measured region starts at ``model(...)``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .ckg import CKG
from .model import KGAT, KGATArgs, KGATMode
from .sampler import BatchSampler

CF_BATCH = 256  # main.py:46 TRAIN_CF_BATCH_SIZE
KG_BATCH = 512  # main.py:47 TRAIN_KG_BATCH_SIZE
CF_LR = 1e-3  # main.py:51
KG_LR = 1e-4  # main.py:52


def attentive_coo(g: CKG) -> torch.Tensor:
    idx = torch.from_numpy(np.vstack([g.att_rows, g.att_cols])).long()
    return torch.sparse_coo_tensor(idx, torch.from_numpy(g.att_vals), size=(g.node_num, g.node_num))


def build_model(g: CKG, device="cuda", seed: int = 2024, **kgat_kwargs) -> KGAT:
    torch.manual_seed(seed)  # main.py:56-67: only torch.manual_seed is effective (Q5)
    model = KGAT(KGATArgs(user_num=g.user_num, entity_num=g.entity_num, relation_num=g.relation_num,
                          attentive_matrix=attentive_coo(g), **kgat_kwargs))
    model.to(device)
    model.build_optimizer(cf_lr=CF_LR, kg_lr=KG_LR)
    return model


@dataclass
class EpochData:
    """One epoch's pre-sampled batches + the refresh edge list (host numpy; ``to`` moves them)."""

    cf: tuple  # (users, pos, neg) each [n_cf, 256] int64
    kg: tuple  # (heads, rels, pos, neg) each [n_kg, 512] int64
    edges: tuple  # (heads int32, rels int64, tails int32, relation_indices int64)

    @property
    def n_cf(self) -> int:
        return int(self.cf[0].shape[0])

    @property
    def n_kg(self) -> int:
        return int(self.kg[0].shape[0])

    @classmethod
    def sample(cls, g: CKG, seed: int = 2024, n_cf: int | None = None, n_kg: int | None = None):
        s = BatchSampler(g, seed)
        n_cf = s.cf_batches_per_epoch(CF_BATCH) if n_cf is None else n_cf
        n_kg = s.kg_batches_per_epoch(KG_BATCH) if n_kg is None else n_kg
        edges = (g.heads, g.relations, g.tails, np.asarray(g.adjacency_relations, np.int64))
        return cls(cf=s.cf_batches(n_cf, CF_BATCH), kg=s.kg_batches(n_kg, KG_BATCH), edges=edges)

    def tensors(self, device=None, pin: bool = False):
        """Torch views of the epoch.  Host tensors keep each step's ids in one [k, B] block (steps stacked as
        [n, k, B]) so a step's inputs travel in ONE host->device copy; ``cf[j][i]`` / ``kg[j][i]`` are views of it."""
        def conv(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if device is not None:
                return t.to(device)
            return t.pin_memory() if pin else t

        out = EpochData(cf=None, kg=None, edges=tuple(conv(a) for a in self.edges))
        out.cf_block = conv(np.stack(self.cf, axis=1))  # [n_cf, 3, B]
        out.kg_block = conv(np.stack(self.kg, axis=1))  # [n_kg, 4, B]
        out.cf = tuple(out.cf_block[:, j] for j in range(3))
        out.kg = tuple(out.kg_block[:, j] for j in range(4))
        return out


def run_epoch(model: KGAT, data: EpochData, read_loss_every_step: bool = False, n_cf: int | None = None, n_kg: int | None = None,
              refresh: bool = True):
    """One reference epoch body through the public model API.

    ``data`` may live on the device (inputs resident in HBM) or in pinned host memory (then every
    step copies its ids host->device, like main.py:302-305 / 329-333).  With
    ``read_loss_every_step`` the loss is read back every step (``.item()``, main.py:314, 343);
    otherwise losses are accumulated on the device and read once at the end.
    Returns (mean CF loss, mean KG loss, bytes host->device, bytes device->host)."""
    dev = model._user_entity_embedding.weight.device
    n_cf = data.n_cf if n_cf is None else min(n_cf, data.n_cf)
    n_kg = data.n_kg if n_kg is None else min(n_kg, data.n_kg)
    h2d = d2h = 0
    model.train()
    cf_sum = torch.zeros((), device=dev)
    cf_host = 0.0
    cf_block, kg_block = getattr(data, "cf_block", None), getattr(data, "kg_block", None)
    for i in range(n_cf):
        if cf_block is not None and not cf_block.is_cuda:
            # one pinned [3, B] block per step, handed to the model as host views: model(...) copies it host->device itself
            # (one cudaMemcpyAsync inside the step's launch call; the API accepts index tensors that live on the CPU)
            u, p, n = cf_block[i].unbind(0)
            h2d += 3 * u.numel() * 8
        else:
            u, p, n = (t[i] for t in data.cf)
            if not u.is_cuda:
                u, p, n = u.to(dev, non_blocking=True), p.to(dev, non_blocking=True), n.to(dev, non_blocking=True)
                h2d += 3 * u.numel() * 8
        loss = model(u, p, n, mode=KGATMode.TRAIN_CF)
        loss.backward()
        model.update_cf_weights()
        if read_loss_every_step:
            cf_host += loss.item()
            d2h += 4
        else:
            cf_sum += loss.detach()
    kg_sum = torch.zeros((), device=dev)
    kg_host = 0.0
    for i in range(n_kg):
        if kg_block is not None and not kg_block.is_cuda:
            h, r, pt, nt = kg_block[i].unbind(0)
            h2d += 4 * h.numel() * 8
        else:
            h, r, pt, nt = (t[i] for t in data.kg)
            if not h.is_cuda:
                h, r, pt, nt = (x.to(dev, non_blocking=True) for x in (h, r, pt, nt))
                h2d += 4 * h.numel() * 8
        loss = model(h, r, pt, nt, mode=KGATMode.TRAIN_KG)
        loss.backward()
        model.update_kg_weights()
        if read_loss_every_step:
            kg_host += loss.item()
            d2h += 4
        else:
            kg_sum += loss.detach()
    if refresh:
        eh, er, et, ri = data.edges
        if not eh.is_cuda:
            h2d += eh.numel() * 4 + er.numel() * 8 + et.numel() * 4 + ri.numel() * 8
            eh, er, et, ri = (x.to(dev, non_blocking=True) for x in (eh, er, et, ri))
        model(eh, er, et, ri, mode=KGATMode.UPDATE_ATTENTION)
    if not read_loss_every_step:
        cf_host, kg_host = float(cf_sum.item()), float(kg_sum.item())
        d2h += 8
    return cf_host / max(n_cf, 1), kg_host / max(n_kg, 1), h2d, d2h
