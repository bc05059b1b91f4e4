#!/usr/bin/env python
"""Short, ncu-friendly slice of the Amazon-book-shaped workload: a few eager (un-captured) CF steps,
KG steps and one attention refresh through the public model API, then the same number of steps through the epoch
engine's un-captured step bodies, so every kernel shows up as its own launch.  Used with the ncu recipes of /opt/skills/guides/B200_PROFILING.md; summaries go to profiles/.

    python tools/prof_steps.py [--shape amazon-book] [--cf 3] [--kg 6]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
from kgat_b200.trainer import build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="amazon-book")
ap.add_argument("--cf", type=int, default=3)
ap.add_argument("--kg", type=int, default=6)
args = ap.parse_args()

g = synthetic.make_ckg(args.shape, with_dicts=False)
model = build_model(g, "cuda").train()
rng = np.random.default_rng(0)
dev = "cuda"
edges = [torch.from_numpy(x).to(dev) for x in (g.heads, g.relations, g.tails, np.asarray(g.adjacency_relations))]
model(*edges, mode=KGATMode.UPDATE_ATTENTION)  # structure build + first refresh (warm-up)
for i in range(args.cf):
    u = torch.from_numpy(rng.choice(g.user_num, 256, replace=False)).to(dev)
    p = torch.from_numpy(rng.integers(0, g.item_num, 256)).to(dev)
    n = torch.from_numpy(rng.integers(0, g.item_num, 256)).to(dev)
    loss = model(u, p, n, mode=KGATMode.TRAIN_CF)
    loss.backward()
    model.update_cf_weights()
for i in range(args.kg):
    sel = rng.integers(0, g.nnz, 512)
    b = [torch.from_numpy(x.astype(np.int64)).to(dev) for x in (g.heads[sel], g.relations[sel], g.tails[sel], rng.integers(0, g.node_num, 512))]
    loss = model(*b, mode=KGATMode.TRAIN_KG)
    loss.backward()
    model.update_kg_weights()
model(*edges, mode=KGATMode.UPDATE_ATTENTION)
with torch.no_grad():
    model.eval()
    s = model(torch.arange(256), torch.arange(g.item_num, device=dev), mode=KGATMode.PREDICT)
# the epoch engine's own step bodies, un-captured (compact KG gradient rows, fused TransR step)
from kgat_b200.engine import TrainEngine  # noqa: E402
from kgat_b200.trainer import EpochData  # noqa: E402

model.train()
g_full = synthetic.make_ckg(args.shape, with_dicts=True)
data = EpochData.sample(g_full, n_cf=args.cf, n_kg=args.kg)
eng = TrainEngine(model, use_graphs=False)
eng.bind_resident(data.tensors())
eng.run_epoch(refresh=False)
torch.cuda.synchronize()
print("ok", float(loss), tuple(s.shape))
