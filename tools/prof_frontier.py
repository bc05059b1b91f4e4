#!/usr/bin/env python
"""Times the level-2 -> level-1 frontier expansion (csrc/frontier.cu) at the Amazon-book shape with CUDA events, one process per
variant (latched at the first call): the shared-memory kernel (default) or KGAT_EXPAND_PLAIN=1.

    python tools/prof_frontier.py; KGAT_EXPAND_PLAIN=1 python tools/prof_frontier.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import ops, synthetic  # noqa: E402
from kgat_b200.frontier import Frontier  # noqa: E402
from kgat_b200.trainer import build_model  # noqa: E402

g = synthetic.make_ckg(sys.argv[1] if len(sys.argv) > 1 else "amazon-book", with_dicts=False)
model = build_model(g, "cuda").train()
graph = model._graph()
rng = np.random.default_rng(0)
dev = "cuda"
f = Frontier(graph, 3, 768)
ids = [torch.from_numpy(rng.choice(g.user_num, 256, replace=False)).to(dev), torch.from_numpy(rng.integers(0, g.item_num, 256)).to(dev),
       torch.from_numpy(rng.integers(0, g.item_num, 256)).to(dev)]  # item ids index the table without a user_num offset (reference quirk Q4)
f.build(ids)
torch.cuda.synchronize()
counts = f.counts.tolist()
ref_mask = f.bitmaps.clone()
reps = 30
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for a, b in ev:
    big.zero_()  # L2 flush
    a.record()
    ops.frontier_expand(graph.plan, graph.col_idx, f.rows(2), f.count(2), f.cap(2), f.mask(2), f.flags)
    b.record()
    ops.frontier_list(f.flags, f.mask(1), f.n, f.scratch, f.rows(1), f.count(1))
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
same = bool((f.bitmaps == ref_mask).all())
print(f"KGAT_EXPAND_PLAIN={os.environ.get('KGAT_EXPAND_PLAIN', '0')} rows L3/L2/L1={counts[2]}/{counts[1]}/{counts[0]} expand(L2->L1) cold-L2 median {t[len(t) // 2]:.1f} us "
      f"min {t[0]:.1f} us  level-1 bitmap unchanged: {same}")
# warm (no flush)
for a, b in ev:
    a.record()
    ops.frontier_expand(graph.plan, graph.col_idx, f.rows(2), f.count(2), f.cap(2), f.mask(2), f.flags)
    b.record()
    ops.frontier_list(f.flags, f.mask(1), f.n, f.scratch, f.rows(1), f.count(1))
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
print(f"   warm median {t[len(t) // 2]:.1f} us min {t[0]:.1f} us")
