"""Debug: lock-step twin models (A: api graphs + pruning, B: eager, no pruning) over CF steps; report first divergence."""
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
data = EpochData.sample(g, seed=5, n_cf=6, n_kg=1).tensors(device="cuda")


def mk(api, prune):
    m = build_model(g, "cuda", seed=5).eval()
    m.api_graphs, m.cf_pruning = api, prune
    return m


for cfgA in ((True, True), (True, False), (False, True)):
    A, B = mk(*cfgA), mk(False, False)
    print("=== A = api %s prune %s" % cfgA)
    for i in range(6):
        ids = [t[i] for t in data.cf]
        la = A(*ids, mode=KGATMode.TRAIN_CF)
        la.backward()
        lb = B(*ids, mode=KGATMode.TRAIN_CF)
        lb.backward()
        torch.cuda.synchronize()
        msg = [f"step {i}: loss A {la.item():.6f} B {lb.item():.6f}"]
        for (k, pa), (_, pb) in zip(A.named_parameters(), B.named_parameters()):
            if pb.grad is None:
                continue
            ga, gb = pa.grad, pb.grad
            fin = bool(torch.isfinite(ga).all())
            err = float((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-30)) if fin else float("nan")
            if (not fin) or err > 1e-4:
                bad_rows = (~torch.isfinite(ga)).reshape(ga.shape[0], -1).any(1).nonzero().flatten() if ga.dim() > 1 else None
                msg.append(f"  grad {k}: finite={fin} err={err:.2e} bad_rows={None if bad_rows is None else bad_rows[:8].tolist()} n_bad={None if bad_rows is None else bad_rows.numel()}")
        A.update_cf_weights()
        B.update_cf_weights()
        torch.cuda.synchronize()
        for (k, pa), (_, pb) in zip(A.named_parameters(), B.named_parameters()):
            if pa.is_sparse:
                continue
            fin = bool(torch.isfinite(pa).all())
            err = float((pa - pb).abs().max())
            if (not fin) or err > 1e-5:
                msg.append(f"  param {k}: finite={fin} maxabs diff={err:.2e}")
        f = A._last_step.get("cf")
        print("\n".join(msg))
