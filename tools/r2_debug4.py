"""Debug: replicate test_engine_and_api_graph_paths[True] api path; locate NaNs."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
from kgat_b200.trainer import EpochData, build_model  # noqa: E402

g = synthetic.make_ckg("small", seed=5)
data = EpochData.sample(g, seed=5, n_cf=3, n_kg=1)
dd = data.tensors(device="cuda")


def unbits(t, n):
    return torch.from_numpy(np.unpackbits(t.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)).cuda()


for trial in range(3):
    model = build_model(g, "cuda", seed=5)
    model.eval()
    model.cf_pruning = True
    model.api_graphs = True
    for i in range(3):
        u, p, n = (t[i] for t in dd.cf)
        loss = model(u, p, n, mode=KGATMode.TRAIN_CF)
        loss.backward()
        torch.cuda.synchronize()
        st = model._last_step["cf"]
        f = st.frontier
        N = g.node_num
        rep = [f"trial {trial} step {i}: item {loss.item():.6f} detach {float(loss.detach().cpu()):.6f} counts {f.counts.cpu().tolist()}"]
        for l, t in enumerate(st.prop.tables):
            lvl = unbits(f.mask(l), N) if l >= 1 else torch.ones(N, dtype=torch.bool, device="cuda")
            bad = (~torch.isfinite(t)).any(1)
            rep.append(f"   table {l}: nan rows inside level {int((bad & lvl).sum())} outside {int((bad & ~lvl).sum())}")
        for l, t in enumerate(st.prop.side):
            lvl = unbits(f.mask(l + 1), N)
            bad = (~torch.isfinite(t)).any(1)
            rep.append(f"   side {l}: nan rows inside level {int((bad & lvl).sum())} outside {int((bad & ~lvl).sum())}")
        gbad = [k for k, q in model.named_parameters() if q.grad is not None and not bool(torch.isfinite(q.grad).all())]
        rep.append(f"   nan grads: {gbad}")
        model.update_cf_weights()
        torch.cuda.synchronize()
        pbad = [k for k, q in model.named_parameters() if not q.is_sparse and not bool(torch.isfinite(q).all())]
        rep.append(f"   nan params after update: {pbad}")
        print("\n".join(rep))
