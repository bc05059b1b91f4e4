"""Debug: KG step through the graphed API path vs the eager autograd path, gradients per replay."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import Golden  # noqa: E402

from kgat_b200.model import KGAT, KGATArgs, KGATMode  # noqa: E402

g = Golden("model_small.npz")


def mk(api):
    m = KGAT(KGATArgs(user_num=int(g["user_num"]), entity_num=int(g["entity_num"]), relation_num=int(g["relation_num"]), attentive_matrix=g.att_coo()))
    m.load_state_dict(g.params(), strict=False)
    m = m.cuda().eval()
    m.api_graphs = api
    m.build_optimizer(cf_lr=1e-3, kg_lr=1e-4)
    return m


kg_b = [torch.from_numpy(g[k]).cuda() for k in ("kg_heads", "kg_rels", "kg_pos", "kg_neg")]
print("ptrs", [t.data_ptr() - kg_b[0].data_ptr() for t in kg_b], "rels", sorted(set(kg_b[1].cpu().tolist())))
for sync in (True, False):
    A, B = mk(True), mk(False)
    for i in range(4):
        la = A(*kg_b, mode=KGATMode.TRAIN_KG)
        la.backward()
        lb = B(*kg_b, mode=KGATMode.TRAIN_KG)
        lb.backward()
        if sync:
            torch.cuda.synchronize()
        rep = []
        for (k, pa), (_, pb) in zip(A.named_parameters(), B.named_parameters()):
            if pb.grad is not None:
                ga, gb = pa.grad.clone(), pb.grad.clone()
                err = float((ga - gb).abs().max() / gb.abs().max())
                rep.append(f"{k}: err {err:.1e} nz rows A {int((ga.reshape(ga.shape[0], -1).abs().sum(1) > 0).sum())} B {int((gb.reshape(gb.shape[0], -1).abs().sum(1) > 0).sum())}")
        A.update_kg_weights()
        B.update_kg_weights()
        pd = {k: f"{float((pa - pb).abs().max()):.1e}" for (k, pa), (_, pb) in zip(A.named_parameters(), B.named_parameters()) if not pa.is_sparse and k in ("_relation_embedding.weight", "_trans_matrix", "_user_entity_embedding.weight")}
        print(f"sync {sync} step {i}: loss {la.item():.6f}/{lb.item():.6f}", rep, "param diff", pd)
