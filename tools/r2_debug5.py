"""Debug: trajectory cf,cf,kg,kg,att,cf on golden small with api graphs + pruning; at the last step compare every
intermediate against an eager unpruned twin with identical parameters."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import Golden  # noqa: E402

from kgat_b200 import functions  # noqa: E402
from kgat_b200.functions import DropoutSpec, propagate_forward  # noqa: E402
from kgat_b200.model import KGAT, KGATArgs, KGATMode  # noqa: E402


def unbits(t, n):
    return torch.from_numpy(np.unpackbits(t.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)).cuda()


for name in ("model_small.npz", "model_tiny.npz"):
    g = Golden(name)
    # recycle some NaN-filled memory first, like a long pytest session would
    junk = [torch.full((2000, 64), float("nan"), device="cuda") for _ in range(40)]
    del junk
    m = KGAT(KGATArgs(user_num=int(g["user_num"]), entity_num=int(g["entity_num"]), relation_num=int(g["relation_num"]), attentive_matrix=g.att_coo()))
    m.load_state_dict(g.params(), strict=False)
    m = m.cuda().eval()
    m.api_graphs, m.cf_pruning = True, True
    m.build_optimizer(cf_lr=1e-3, kg_lr=1e-4)
    cf_b = [torch.from_numpy(g[k]).cuda() for k in ("cf_users", "cf_pos", "cf_neg")]
    kg_b = [torch.from_numpy(g[k]).cuda() for k in ("kg_heads", "kg_rels", "kg_pos", "kg_neg")]
    N = g.node_num
    for what in ("cf", "cf", "kg", "kg", "att", "cf"):
        if what == "cf":
            sd = {k: v.detach().clone() for k, v in m.state_dict().items() if not v.is_sparse}
            loss = m(*cf_b, mode=KGATMode.TRAIN_CF)
            loss.backward()
            torch.cuda.synchronize()
            st = m._last_step["cf"]
            f = st.frontier
            graph = m._graph()
            layers = [tuple(t.detach() for t in grp) for grp in m._layers()]
            ref = propagate_forward(graph, sd["_user_entity_embedding.weight"], [tuple(sd[f"_aggregator_layers.{l}.{w}"] for w in ("linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias")) for l in range(3)], DropoutSpec(ps=[0.0] * 3), save=True)
            print(name, what, "loss", loss.item(), "counts", f.counts.cpu().tolist(), "N", N, "batch", cf_b[0].numel())
            for l in range(1, 4):
                lvl = unbits(f.mask(l), N)
                rows = f.rows(l)[: int(f.count(l).item())].long()
                assert bool((lvl.nonzero().flatten() == rows).all())
                a, b = st.prop.tables[l], ref.tables[l]
                d = (a - b).abs().max(1).values
                bad_in = ((d > 1e-5) | ~torch.isfinite(d)) & lvl
                sa, sb = st.prop.side[l - 1], ref.side[l - 1]
                ds = (sa - sb).abs().max(1).values
                bad_s = ((ds > 1e-5) | ~torch.isfinite(ds)) & lvl
                print(f"   level {l}: rows {rows.numel()} table bad rows inside {int(bad_in.sum())} {bad_in.nonzero().flatten()[:6].tolist()} side bad rows inside {int(bad_s.sum())} {bad_s.nonzero().flatten()[:6].tolist()}")
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
