#!/usr/bin/env python
"""KG phase of the Amazon-book-shaped workload through the epoch engine (captured graphs), for ncu launch lists and A/B timing of
the KG-phase optimiser (KGAT_KG_ADAM = rolling | dense | lazy, KGAT_KG_WINDOW).

    python tools/prof_kg.py [--kg 600] [--epochs 2]
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
ap.add_argument("--kg", type=int, default=600)
ap.add_argument("--epochs", type=int, default=2)
args = ap.parse_args()

g = synthetic.make_ckg(args.shape, with_dicts=True)
model = build_model(g, "cuda").train()
data = EpochData.sample(g, n_cf=1, n_kg=args.kg)
eng = TrainEngine(model, kg_window=int(os.environ.get("KGAT_KG_WINDOW", "16")))
eng.bind_resident(data.tensors())
for e in range(args.epochs):
    eng.run_epoch(n_cf=0, refresh=False)
    torch.cuda.synchronize()
    print(f"mode={eng.kg_adam_mode} window={eng.kg_window} epoch {e}: kg_step_us = {1e3 * eng.last_phase_ms['kg'] / args.kg:.2f}")
