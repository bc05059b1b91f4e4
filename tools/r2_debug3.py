"""Debug: api graphs + pruning: (1) with every buffer re-poisoned each replay; (2) Adam plan on/off."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from kgat_b200 import functions, synthetic  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
from kgat_b200.trainer import EpochData, build_model  # noqa: E402

g = synthetic.make_ckg("small", seed=5)
data = EpochData.sample(g, seed=5, n_cf=6, n_kg=1).tensors(device="cuda")


def mk(api, prune):
    m = build_model(g, "cuda", seed=5).eval()
    m.api_graphs, m.cf_pruning = api, prune
    return m


for poison, adam_graphs in ((True, True), (False, False), (False, True)):
    functions.POISON_STALE_ROWS = poison
    A, B = mk(True, True), mk(False, False)
    A._cf_optimizer.use_graphs = adam_graphs
    print(f"=== poison {poison} adam_graphs {adam_graphs}")
    for i in range(5):
        ids = [t[i] for t in data.cf]
        la = A(*ids, mode=KGATMode.TRAIN_CF)
        la.backward()
        functions.POISON_STALE_ROWS = False
        lb = B(*ids, mode=KGATMode.TRAIN_CF)
        lb.backward()
        functions.POISON_STALE_ROWS = poison
        torch.cuda.synchronize()
        snap = {k: p.grad.clone() for k, p in A.named_parameters() if p.grad is not None}
        worst = max(float((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp_min(1e-30)) for (k, p), (_, q) in zip(A.named_parameters(), B.named_parameters()) if q.grad is not None)
        fin = all(bool(torch.isfinite(v).all()) for v in snap.values())
        st = A._last_step["cf"]
        static = {k: gr for (k, p), gr in zip([(k, p) for k, p in A.named_parameters() if p.grad is not None], st.grads)}
        A.update_cf_weights()
        B.update_cf_weights()
        torch.cuda.synchronize()
        changed = [k for k in snap if not torch.equal(snap[k], static[k])]
        pdiff = {k: float((p - q).abs().max()) for (k, p), (_, q) in zip(A.named_parameters(), B.named_parameters()) if not p.is_sparse}
        big = {k: f"{v:.1e}" for k, v in pdiff.items() if v > 2e-5}
        print(f"step {i}: loss A {la.item():.6f} B {lb.item():.6f} grads finite {fin} worst grad err {worst:.1e} static grads changed by update: {changed} param diffs: {big}")
functions.POISON_STALE_ROWS = False
