"""Debug: full trajectory on the goldens: parameters after the run vs the reference's, for api graphs / pruning on/off."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import Golden  # noqa: E402

from kgat_b200.model import KGAT, KGATArgs, KGATMode  # noqa: E402

for name in ("model_tiny.npz", "model_small.npz"):
    g = Golden(name)
    for api, prune, adam in ((True, True, True), (False, False, False), (True, False, True), (False, True, True), (True, True, False)):
        m = KGAT(KGATArgs(user_num=int(g["user_num"]), entity_num=int(g["entity_num"]), relation_num=int(g["relation_num"]), attentive_matrix=g.att_coo()))
        m.load_state_dict(g.params(), strict=False)
        m = m.cuda().eval()
        m.api_graphs, m.cf_pruning = api, prune
        m.build_optimizer(cf_lr=1e-3, kg_lr=1e-4)
        m._cf_optimizer.use_graphs = m._kg_optimizer.use_graphs = adam
        cf_b = [torch.from_numpy(g[k]).cuda() for k in ("cf_users", "cf_pos", "cf_neg")]
        kg_b = [torch.from_numpy(g[k]).cuda() for k in ("kg_heads", "kg_rels", "kg_pos", "kg_neg")]
        rel_hist = []
        for what in ("cf", "cf", "kg", "kg", "att", "cf"):
            before = m._relation_embedding.weight.detach().clone()
            if what == "cf":
                loss = m(*cf_b, mode=KGATMode.TRAIN_CF)
                loss.backward()
                m.update_cf_weights()
            elif what == "kg":
                loss = m(*kg_b, mode=KGATMode.TRAIN_KG)
                loss.backward()
                m.update_kg_weights()
            else:
                heads = torch.tensor(list(g["heads"].astype(np.int32))).cuda()
                rels = torch.tensor(g["relations"].tolist()).cuda()
                tails = torch.tensor(list(g["tails"].astype(np.int32))).cuda()
                m(heads, rels, tails, torch.tensor(g["adjacency_relations"].tolist()).cuda(), mode=KGATMode.UPDATE_ATTENTION)
            rel_hist.append(f"{what}:{float((m._relation_embedding.weight.detach() - before).abs().mean()):.2e}")
        sd = m.state_dict()
        diffs = {}
        for k in g.keys():
            if k.startswith("traj_param::"):
                got, ref = sd[k[12:]].cpu().double(), torch.from_numpy(g[k]).double()
                diffs[k[12:].replace("_aggregator_layers", "agg")] = f"{float((got - ref).abs().mean()):.1e}"
        print(name, "api", api, "prune", prune, "adam plans", adam, "| rel-emb movement per step", rel_hist)
        print("      mean |param - ref|:", diffs)
