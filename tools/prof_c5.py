#!/usr/bin/env python
"""The HBM-regime SpMM of bench.py's ``c5_scaled`` block (configs[4] scaled 5x down: 2.2 M nodes, 40 M edges, d = 128; the
1.1 GB table is 9x the L2) as a stand-alone launch sequence for ncu:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
        --clock-control none -k regex:spmm_task --csv --log-file out.csv python tools/prof_c5.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.graph import AttentiveGraph  # noqa: E402

dev = "cuda"
n5, d5 = 2_200_000, 128
h5, _, t5 = synthetic.make_edges_only(n5, 40_000_000, 64)
deg5 = np.bincount(h5, minlength=n5).astype(np.float32)
g5 = AttentiveGraph.from_coo(torch.from_numpy(h5.astype(np.int64)).to(dev), torch.from_numpy(t5.astype(np.int64)).to(dev),
                             torch.from_numpy((1.0 / deg5[h5]).astype(np.float32)).to(dev), n5)
torch.manual_seed(0)
x = torch.randn(n5, d5, device=dev) * 0.1
y = torch.empty_like(x)
for _ in range(4):
    g5.matmul(x, out=y)
torch.cuda.synchronize()
b_gather = 8.0 * g5.nnz + 16.0 * g5.plan.n_tasks + 4.0 * g5.nnz * d5 + 4.0 * n5 * d5
b_min = 8.0 * g5.nnz + 16.0 * g5.plan.n_tasks + 2 * 4.0 * n5 * d5
print(f"nnz={g5.nnz} tasks={g5.plan.n_tasks} B_gather={b_gather / 1e9:.3f} GB B_min={b_min / 1e9:.3f} GB")
