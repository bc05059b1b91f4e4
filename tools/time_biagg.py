#!/usr/bin/env python
"""CUDA-event timing of the bi-interaction forward / backward kernels at the C3 row count, plus accuracy against an fp64
torch evaluation.  Run once per implementation: KGAT_BIAGG_IMPL=ffma|mma|tc5 python tools/time_biagg.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from kgat_b200 import ops  # noqa: E402

n = int(os.environ.get("KGAT_TIME_ROWS", "159251"))
torch.manual_seed(0)
print("impl", os.environ.get("KGAT_BIAGG_IMPL", "(default)"), "rows", n)
for d_in, d_out in ((64, 64), (64, 32), (32, 16)):
    E, S = torch.randn(n, d_in, device="cuda"), torch.randn(n, d_in, device="cuda")
    W1, W2 = torch.randn(d_out, d_in, device="cuda") / 8, torch.randn(d_out, d_in, device="cuda") / 8
    b1, b2 = torch.randn(d_out, device="cuda") * 0.1, torch.randn(d_out, device="cuda") * 0.1
    out = torch.empty(n, d_out, device="cuda")
    inv = torch.empty(n, device="cuda")
    flags = torch.empty(n, d_out, dtype=torch.uint8, device="cuda")
    g = torch.randn(n, d_out, device="cuda")
    nc = ops.biagg_backward_ctas(n, d_in, d_out)
    part = torch.empty(nc * (2 * d_in * d_out + 2 * d_out), device="cuda")
    gs, ge = torch.empty_like(E), torch.empty_like(E)
    flush = torch.empty(64 << 20, device="cuda")

    def timed(fn, reps=20):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    fwd = lambda: ops.biagg_forward(E, S, W1, b1, W2, b2, out, inv, flags, dropout_p=0.1, seed=1, offset=0)  # noqa: E731
    bwd = lambda: ops.biagg_backward(g, out, inv, flags, E, S, W1, W2, 0.1, gs, ge, part, nc)  # noqa: E731
    for _ in range(3):
        fwd(), bwd()
    t_f, t_b = timed(fwd), timed(bwd)
    # accuracy without dropout, fp64 reference
    ops.biagg_forward(E, S, W1, b1, W2, b2, out, inv, flags, dropout_p=0.0)
    Ed, Sd = E.double(), S.double()
    lre = torch.nn.functional.leaky_relu
    x = lre((Ed + Sd) @ W1.double().T + b1.double(), 0.01) + lre((Ed * Sd) @ W2.double().T + b2.double(), 0.01)
    ref = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    err = float((out.double() - ref).norm() / ref.norm())
    print(f"{d_in:3d}->{d_out:3d}  fwd {t_f:7.1f} us   bwd {t_b:7.1f} us   fwd rel err vs fp64 {err:.2e}")
