"""Debug: the reference trajectory (cf, cf, kg, kg, att, cf) under every combination of api_graphs x cf_pruning."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import Golden  # noqa: E402

from kgat_b200.model import KGAT, KGATArgs, KGATMode  # noqa: E402

import os
NAMES = os.environ.get("DBG_NAMES", "model_tiny.npz,model_small.npz").split(",")
CFGS = [tuple(c == "T" for c in x) for x in os.environ.get("DBG_CFGS", "FF,FT,TF,TT").split(",")]
for name in NAMES:
    g = Golden(name)
    for api, prune in CFGS:
        if True:
            m = KGAT(KGATArgs(user_num=int(g["user_num"]), entity_num=int(g["entity_num"]), relation_num=int(g["relation_num"]), attentive_matrix=g.att_coo()))
            m.load_state_dict(g.params(), strict=False)
            m = m.cuda().eval()
            m.api_graphs, m.cf_pruning = api, prune
            m.build_optimizer(cf_lr=1e-3, kg_lr=1e-4)
            cf_b = [torch.from_numpy(g[k]).cuda() for k in ("cf_users", "cf_pos", "cf_neg")]
            kg_b = [torch.from_numpy(g[k]).cuda() for k in ("kg_heads", "kg_rels", "kg_pos", "kg_neg")]
            losses = []
            for what in ("cf", "cf", "kg", "kg", "att", "cf", "cf"):
                if what == "cf":
                    loss = m(*cf_b, mode=KGATMode.TRAIN_CF)
                    loss.backward()
                    bad = [k for k, p in m.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
                    m.update_cf_weights()
                    losses.append((loss.item(), bad))
                elif what == "kg":
                    loss = m(*kg_b, mode=KGATMode.TRAIN_KG)
                    loss.backward()
                    m.update_kg_weights()
                    losses.append((loss.item(), []))
                else:
                    heads = torch.tensor(list(g["heads"].astype(np.int32))).cuda()
                    rels = torch.tensor(g["relations"].tolist()).cuda()
                    tails = torch.tensor(list(g["tails"].astype(np.int32))).cuda()
                    m(heads, rels, tails, torch.tensor(g["adjacency_relations"].tolist()).cuda(), mode=KGATMode.UPDATE_ATTENTION)
                    gr = m._graph()
                    print("   after refresh: vals finite", bool(torch.isfinite(gr.vals).all()), "t_vals finite", bool(torch.isfinite(gr.t_vals).all()),
                          "n_heavy", gr.plan.n_heavy, gr.t_plan.n_heavy)
            print(name, "api", api, "prune", prune, [(round(l, 6), b) for l, b in losses], "ref", np.round(g["traj_losses"], 6).tolist())
