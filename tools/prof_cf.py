#!/usr/bin/env python
"""CF phase of the Amazon-book-shaped workload through the epoch engine (captured graphs): per-step time for A/B runs of kernel
variants selected by environment switches (KGAT_SPMM_HALF, KGAT_SPMM_U, ...), one process per variant.

    python tools/prof_cf.py [--cf 300] [--epochs 3]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.engine import TrainEngine  # noqa: E402
from kgat_b200.trainer import EpochData, build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="amazon-book")
ap.add_argument("--cf", type=int, default=300)
ap.add_argument("--epochs", type=int, default=3)
args = ap.parse_args()

g = synthetic.make_ckg(args.shape, with_dicts=True)
model = build_model(g, "cuda").train()
data = EpochData.sample(g, n_cf=args.cf, n_kg=1)
eng = TrainEngine(model)
eng.bind_resident(data.tensors())
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("KGAT_"))
for e in range(args.epochs):
    loss = eng.run_epoch(n_kg=0, refresh=False)[0]
    torch.cuda.synchronize()
    print(f"[{tag}] epoch {e}: cf_step_us = {1e3 * eng.last_phase_ms['cf'] / args.cf:.1f}  loss {loss:.6f}")
