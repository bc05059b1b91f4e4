#!/usr/bin/env python
"""ncu target: the bi-interaction forward / backward kernels alone at the C3 row count."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from kgat_b200 import ops  # noqa: E402

n = 159251
torch.manual_seed(0)
for d_in, d_out in ((64, 64), (64, 32)):
    E, S = torch.randn(n, d_in, device="cuda"), torch.randn(n, d_in, device="cuda")
    W1, W2 = torch.randn(d_out, d_in, device="cuda") / 8, torch.randn(d_out, d_in, device="cuda") / 8
    b1, b2 = torch.zeros(d_out, device="cuda"), torch.zeros(d_out, device="cuda")
    out = torch.empty(n, d_out, device="cuda")
    inv = torch.empty(n, device="cuda")
    flags = torch.empty(n, d_out, dtype=torch.uint8, device="cuda")
    g = torch.randn(n, d_out, device="cuda")
    nc = ops.biagg_backward_ctas(n, d_in, d_out)
    part = torch.empty(nc * (2 * d_in * d_out + 2 * d_out), device="cuda")
    gs, ge = torch.empty_like(E), torch.empty_like(E)
    for _ in range(3):
        ops.biagg_forward(E, S, W1, b1, W2, b2, out, inv, flags, dropout_p=0.1, seed=1, offset=0)
        ops.biagg_backward(g, out, inv, flags, E, S, W1, W2, 0.1, gs, ge, part, nc)
torch.cuda.synchronize()
print("ok")
