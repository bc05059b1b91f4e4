#!/usr/bin/env python
"""Host-side cost of the reference-facing model API: cProfile over KG / CF training steps driven exactly like
bench.py's e2e leg (trainer.run_epoch with pinned host batches).  The KG step is ~80 us of GPU work, so the
API path is bound by Python; this shows where.

    python tools/prof_api.py [--shape amazon-book] [--cf 300] [--kg 3000]
"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.trainer import EpochData, build_model, run_epoch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="amazon-book")
ap.add_argument("--cf", type=int, default=300)
ap.add_argument("--kg", type=int, default=3000)
ap.add_argument("--top", type=int, default=45)
args = ap.parse_args()

g = synthetic.make_ckg(args.shape)
model = build_model(g, "cuda").train()
data = EpochData.sample(g, n_cf=args.cf, n_kg=args.kg).tensors(pin=True)
for _ in range(2):
    run_epoch(model, data, n_cf=20, n_kg=50, refresh=True)
torch.cuda.synchronize()
for name, kw in (("kg", dict(n_cf=0, n_kg=args.kg)), ("cf", dict(n_cf=args.cf, n_kg=0))):
    t0 = time.perf_counter()
    run_epoch(model, data, refresh=False, read_loss_every_step=True, **kw)
    torch.cuda.synchronize()
    n = args.kg if name == "kg" else args.cf
    print(f"{name}: {1e6 * (time.perf_counter() - t0) / n:.1f} us/step wall (unprofiled)")
    pr = cProfile.Profile()
    pr.enable()
    run_epoch(model, data, refresh=False, read_loss_every_step=True, **kw)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(args.top)
