#!/usr/bin/env python
"""Benchmark of the KGAT hot path on the Amazon-book-shaped synthetic CKG (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W             # our arm (CUDA kernels)
    python bench.py --impl reference --steps K --warmup W     # reference arm: CPU oracle port

Metric (BASELINE.json): **KGAT epoch seconds** -- one "step" is one reference epoch body
(main.py:290-361: n_cf CF steps + n_kg KG steps + one attention refresh) over pre-sampled synthetic
batches; ``value`` is measured with the batches resident in HBM, ``e2e`` through the same public
model API with the batches in pinned host memory, a host->device copy of every step's ids and a
device->host read of every step's loss inside the timed region.  Extra keys report propagation
edges/s, the per-kernel time split, the roofline of the dominant kernel (live CUDA-event timing) and
the CPU baseline (the oracle port of the reference's PyTorch-CPU path, timed on this box's cores on a
bounded sample and extrapolated to one epoch).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "kgat_epoch_seconds"
UNIT = "s"
WORKLOAD = "amazon-book"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int = 0):
        self.rows: list[list[str]] = []
        self.proc = None
        self.index = index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's PyTorch-CPU path
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(g, params, data, n_cf_steps: int = 1, n_kg_steps: int = 3, refresh_relations: int = 2):
    """Times a bounded sample of the epoch on the host cores with the oracle (a line-by-line port of
    the reference's ATen call sequence: sparse-COO matmul, nn.Linear, bmm, CPU sparse softmax) and
    extrapolates to the full epoch with the reference's batch counts.  Returns a dict."""
    from oracle import kgat_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = g.node_num
    att = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([g.att_rows, g.att_cols])).long(), torch.from_numpy(g.att_vals), size=(n, n))
    p = {k: v.detach().clone().cpu() for k, v in params.items()}
    states = {"cf": {}, "kg": {}}

    def step(kind, loss_fn, lr, step_no):
        leaves = {k: v.requires_grad_(True) for k, v in p.items()}
        t0 = time.perf_counter()
        loss = loss_fn(leaves)
        loss.backward()
        with torch.no_grad():
            for k, leaf in leaves.items():
                if leaf.grad is None:
                    continue
                m, v = states[kind].setdefault(k, (torch.zeros_like(leaf), torch.zeros_like(leaf)))
                O.adam_step(leaf, leaf.grad, m, v, step_no, lr)
                leaf.grad = None
        float(loss)
        return time.perf_counter() - t0

    cf_t = []
    for i in range(n_cf_steps):
        b = [torch.from_numpy(a[i]) for a in data.cf]
        cf_t.append(step("cf", lambda q: O.cf_loss(q, att, *b), 1e-3, i + 1))
    kg_t = []
    for i in range(n_kg_steps):
        b = [torch.from_numpy(a[i]) for a in data.kg]
        kg_t.append(step("kg", lambda q: O.kg_loss(q, *b), 1e-4, i + 1))
    # attention refresh on the edges of the first `refresh_relations` relation ids, scaled by edge share
    rels = np.asarray(g.adjacency_relations[:refresh_relations])
    sel = np.isin(g.relations, rels)
    t0 = time.perf_counter()
    with torch.no_grad():
        pp = {k: v.detach() for k, v in p.items()}
        O.attention_refresh(pp, g.heads[sel].astype(np.int64), g.relations[sel], g.tails[sel].astype(np.int64), rels, n)
    t_ref_part = time.perf_counter() - t0
    share = float(sel.sum()) / max(g.nnz, 1)
    t_refresh = t_ref_part / max(share, 1e-9)
    t_cf, t_kg = float(np.mean(cf_t)), float(np.mean(kg_t))
    epoch = data.n_cf * t_cf + data.n_kg * t_kg + t_refresh
    return {
        "epoch_s_extrapolated": epoch, "cf_step_s": t_cf, "kg_step_s": t_kg, "refresh_s_extrapolated": t_refresh,
        "cores": threads,
        "sample": f"{n_cf_steps} CF step(s) + {n_kg_steps} KG steps + refresh of {share:.1%} of the edges, extrapolated x({data.n_cf}, {data.n_kg}, 1/share)",
        "sample_wall_s": float(sum(cf_t) + sum(kg_t) + t_ref_part),
    }


def aten_gpu_sample(g, params, data, device, n_cf_steps: int = 3, n_kg_steps: int = 10, refresh_relations: int = 2):
    """The same oracle port (the reference's ATen call sequence) with every tensor on the B200: what stock PyTorch
    -- cuSPARSE SpMM, cuBLAS, ~100 elementwise kernels per step, torch.optim.Adam, the device -> host -> device round
    trip around the CPU sparse softmax of model.py:364-366 -- makes of this epoch on the same box (SURVEY.md 8d:
    "the real kernel to beat").  Part of the baseline leg: bounded sample, extrapolated like the CPU one."""
    from oracle import kgat_oracle as O

    n = g.node_num
    att = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([g.att_rows, g.att_cols])).long(), torch.from_numpy(g.att_vals),
                                  size=(n, n)).to(device)
    p = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in params.items()}
    opt = {"cf": torch.optim.Adam(list(p.values()), lr=1e-3), "kg": torch.optim.Adam(list(p.values()), lr=1e-4)}

    def run(kind, batches, loss_fn):
        ts = []
        for i, b in enumerate(batches):
            b = [torch.from_numpy(a).to(device) for a in b]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss = loss_fn(p, b)
            loss.backward()
            opt[kind].step()
            opt[kind].zero_grad()
            float(loss)  # main.py:314 / 343 read the loss every step
            ts.append(time.perf_counter() - t0)
        return float(np.mean(ts[1:]))  # first step = warm-up (cuSPARSE / cuBLAS handles, allocator)

    t_cf = run("cf", [[a[i] for a in data.cf] for i in range(n_cf_steps + 1)], lambda q, b: O.cf_loss(q, att, *b))
    t_kg = run("kg", [[a[i] for a in data.kg] for i in range(n_kg_steps + 1)], lambda q, b: O.kg_loss(q, *b))
    rels = np.asarray(g.adjacency_relations[:refresh_relations])
    sel = np.isin(g.relations, rels)
    heads = torch.from_numpy(g.heads[sel].astype(np.int64)).to(device)
    tails = torch.from_numpy(g.tails[sel].astype(np.int64)).to(device)
    rel_of = torch.from_numpy(g.relations[sel].astype(np.int64)).to(device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        q = {k: v.detach() for k, v in p.items()}
        rows, cols, vals = [], [], []
        for r in rels.tolist():  # model.py:342-353
            idx = torch.where(rel_of == r)[0]
            rows.append(heads[idx])
            cols.append(tails[idx])
            vals.append(O.attention_by_relation(q, heads[idx], tails[idx], r, n))
        m = torch.sparse_coo_tensor(torch.stack([torch.cat(rows), torch.cat(cols)]), torch.cat(vals), size=(n, n))
        m = torch.sparse.softmax(m.cpu(), dim=1).to(device)  # model.py:364-366
    torch.cuda.synchronize()
    share = float(sel.sum()) / max(g.nnz, 1)
    t_refresh = (time.perf_counter() - t0) / max(share, 1e-9)
    return {"value": data.n_cf * t_cf + data.n_kg * t_kg + t_refresh, "unit": UNIT,
            "kind": "oracle port with all tensors on cuda:0 (stock ATen / cuSPARSE / torch.optim.Adam kernels, loss read every step)",
            "cf_step_s": t_cf, "kg_step_s": t_kg, "refresh_s": t_refresh,
            "sample": f"{n_cf_steps} CF + {n_kg_steps} KG steps after one warm-up step each + refresh of {share:.1%} of the edges, extrapolated"}


def make_workload(workload: str, seed: int = 2024):
    from kgat_b200 import synthetic
    from kgat_b200.trainer import EpochData

    g = synthetic.make_ckg(workload, seed=seed, with_dicts=False)
    data = EpochData.sample(g, seed=seed)
    return g, data


def init_params(g, seed: int = 2024):
    """Reference-initialised parameters (KGAT module constructed on the CPU under the seed)."""
    from kgat_b200.model import KGAT, KGATArgs

    torch.manual_seed(seed)
    m = KGAT(KGATArgs(user_num=g.user_num, entity_num=g.entity_num, relation_num=g.relation_num))
    return {k: v.detach().clone() for k, v in m.state_dict().items() if not v.is_sparse}


def config_dict(g, data, n_gpus):
    return {
        "workload": f"synthetic {WORKLOAD}-shaped CKG (configs[2]): users={g.user_num} items={g.item_num} entities={g.entity_num} "
                    f"relations={g.relation_num} nodes={g.node_num} nnz={g.nnz}; 3 layers 64->64->32->16, d=64, fp32",
        "epoch": f"{data.n_cf} CF steps (B=256) + {data.n_kg} KG steps (B=512) + 1 attention refresh",
        "l2_policy": "inputs larger than L2: the epoch streams >1 TB through 126 MB of L2; every step rewrites the 41 MB table, its Adam moments and ~0.5 GB of activations",
        "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, data = make_workload(WORKLOAD)
    params = init_params(g)
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(g, params, data)
        if i >= args.warmup:
            vals.append(last["epoch_s_extrapolated"])
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(g, data, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "detail": {k: last[k] for k in ("cf_step_s", "kg_step_s", "refresh_s_extrapolated")},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(name: str, g, graph) -> float | None:
    """Compulsory-traffic model B_min of SURVEY.md section 8d, per launch of the named kernel."""
    n, nnz = g.node_num, graph.nnz
    name = name.replace("_pruned", "")
    if name.startswith("spmm"):
        d = int(name.split("_d")[1])
        transposed = name.startswith("spmmT")
        plan = graph.t_plan if transposed else graph.plan
        b = 8.0 * nnz + 16.0 * plan.n_tasks + 4.0 * n * d * 2  # (col, val) + task list + read X once + write Y
        if transposed:
            b += 4.0 * n * d  # addend (direct gradient) read
        return b
    if name.startswith("biagg_fwd_"):
        di, do = (int(x) for x in name[len("biagg_fwd_"):].split("x"))
        return n * (4.0 * (2 * di + do) + 4 + do)
    if name.startswith("biagg_bwd_"):
        di, do = (int(x) for x in name[len("biagg_bwd_"):].split("x"))
        return n * (4.0 * (2 * do + 2 * di + 2 * di) + 4 + do)
    if name == "adam_apply":
        return None  # depends on the tensor set; reported separately
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hbm-regime", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import kgat_b200  # noqa: F401
    from kgat_b200 import _lib, ops
    from kgat_b200.engine import TrainEngine
    from kgat_b200.trainer import build_model, run_epoch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        if args.gpus != world:
            raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
        from kgat_b200 import sharding

        sharding.bench_main(args, METRIC, UNIT, WORKLOAD, make_workload, config_dict, ClockSampler)
        return
    if args.gpus != 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    g, data = make_workload(WORKLOAD)
    model = build_model(g, dev)
    init_state = {k: v.detach().clone().cpu() for k, v in model.state_dict().items() if not v.is_sparse}
    dev_data = data.tensors(device=dev)
    host_data = data.tensors(pin=True)

    def barrier():
        torch.cuda.synchronize()

    engine = TrainEngine(model)
    engine.bind_resident(data.tensors())

    # ---- set-up (never timed): two 2-step mini-epochs capture the step graphs and take the one-off structure swap of
    #      the first attention refresh, so that even --warmup 0 times steady-state epochs only
    for _ in range(2):
        engine.run_epoch(n_cf=2, n_kg=2)
    # ---- warm-up: W full epochs (>= 3 by contract) through the CUDA-graph engine ----
    for _ in range(max(args.warmup, 0)):
        engine.run_epoch()
    barrier()

    # ---- timed: K epochs, inputs resident in HBM ----
    _lib.LaunchCounter.count = 0
    with ClockSampler(local_rank) as clocks:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        losses = None
        for _ in range(args.steps):
            losses = engine.run_epoch()
        t1.record()
        barrier()
    launches = _lib.LaunchCounter.count
    phases = dict(engine.last_phase_ms)
    epoch_s = t0.elapsed_time(t1) / 1e3 / max(args.steps, 1)
    clk = clocks.summary()

    # ---- e2e: one epoch through the reference-facing model API (model(...), loss.backward(),
    #      update_*_weights(), loss.item()) from pinned host buffers; and the same through the engine ----
    e2e = None
    if not args.no_e2e:
        steps_in_epoch = data.n_cf + data.n_kg
        barrier()
        w0 = time.perf_counter()
        _, _, h2d, d2h = run_epoch(model, host_data, read_loss_every_step=True)
        barrier()
        e2e_s = time.perf_counter() - w0
        engine = TrainEngine(model)  # re-read the Adam step counts the API epoch advanced
        engine.bind_resident(data.tensors())
        engine.run_epoch(host_data, read_loss_every_step=True, n_cf=8, n_kg=8, refresh=False)  # capture host-mode graphs
        barrier()
        w0 = time.perf_counter()
        engine.run_epoch(host_data, read_loss_every_step=True)
        barrier()
        e2e_engine_s = time.perf_counter() - w0
        e2e = {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "model(..., mode=TRAIN_CF/TRAIN_KG/UPDATE_ATTENTION) + loss.backward() + update_*_weights() + loss.item()",
               "engine_value": e2e_engine_s,
               "note": f"bytes are per epoch (= one bench step: {steps_in_epoch} model steps; ids copied host->device and the loss read back every model step); "
                       "engine_value = same host buffers through the CUDA-graph TrainEngine"}

    # ---- per-kernel split + roofline of the dominant kernel (live CUDA events, same process) ----
    n_probe_cf, n_probe_kg = 30, 60
    # eager (un-captured) launches so that every kernel can be bracketed by its own pair of CUDA events
    model.api_graphs = False
    model._cf_optimizer.use_graphs = model._kg_optimizer.use_graphs = False
    with ops.KernelTimer() as kt:
        run_epoch(model, dev_data, n_cf=n_probe_cf, n_kg=n_probe_kg, refresh=False)
    model.api_graphs = True
    model._cf_optimizer.use_graphs = model._kg_optimizer.use_graphs = True
    ksum = kt.summary()
    per_epoch_ms = {}
    for name, (cnt, ms) in ksum.items():
        cf_kernel = not name.startswith("transr")
        if name == "adam_apply":
            continue
        scale = (data.n_cf / n_probe_cf) if cf_kernel else (data.n_kg / n_probe_kg)
        per_epoch_ms[name] = {"launches_per_epoch": int(cnt * scale), "avg_us": 1e3 * ms / cnt, "epoch_ms": ms * scale}
    # adam_apply runs in both phases with different tensor sets: split by position
    if "adam_apply" in kt.events:
        ev = kt.events["adam_apply"]
        cf_ms = sum(a.elapsed_time(b) for a, b in ev[:n_probe_cf])
        kg_ms = sum(a.elapsed_time(b) for a, b in ev[n_probe_cf:])
        per_epoch_ms["adam_apply_cf"] = {"launches_per_epoch": data.n_cf, "avg_us": 1e3 * cf_ms / n_probe_cf, "epoch_ms": cf_ms * data.n_cf / n_probe_cf}
        per_epoch_ms["adam_apply_kg"] = {"launches_per_epoch": data.n_kg, "avg_us": 1e3 * kg_ms / max(n_probe_kg, 1), "epoch_ms": kg_ms * data.n_kg / max(n_probe_kg, 1)}
    # dominant kernel = the propagation kernel (north_star) with the largest share of the epoch; the eager probe inflates
    # the ~50 us kernels of the KG phase (launch gaps between the two events), so they are reported separately below
    top = max((k for k in per_epoch_ms if k.startswith("spmm")), key=lambda k: per_epoch_ms[k]["epoch_ms"])
    peak, peak_kind = peaks()
    graph = model._graph()
    emb_numel = model._user_entity_embedding.weight.numel()
    if top.startswith("adam_apply"):
        abytes = 7.0 * 4 * emb_numel
    else:
        abytes = algorithmic_bytes(top, g, graph)
    achieved = abytes / (per_epoch_ms[top]["avg_us"] * 1e-6) / 1e9 if abytes else None
    traffic = None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(top)
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs, burst copy)", "algorithmic_bytes_per_launch": abytes,
                "model": "B_min (every distinct byte once, SURVEY.md 8d); gather_gbs adds one neighbour-row read per edge (L2 traffic)"}
    if top.startswith("spmm"):
        d = int(top.replace("_pruned", "").split("_d")[1])
        roofline["gather_gbs"] = (abytes + 4.0 * graph.nnz * d) / (per_epoch_ms[top]["avg_us"] * 1e-6) / 1e9
    # the dense Adam sweep of the KG phase (second-largest single kernel), timed back to back so launch gaps do not count
    ad = engine.kg_adam
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    snap = ad.snapshot()
    e0.record()
    for _ in range(20):
        ad.apply(engine.kg_grads)
    e1.record()
    torch.cuda.synchronize()
    ad.restore(snap)
    adam_us = 1e3 * e0.elapsed_time(e1) / 20
    adam_bytes = 7.0 * 4 * sum(p.numel() for p in ad.params)
    roofline_adam = {"kernel": "adam_kernel (KG phase, dense sweep)", "bound": "hbm", "avg_us": adam_us, "algorithmic_bytes_per_launch": adam_bytes,
                     "achieved": adam_bytes / (adam_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s", "frac": adam_bytes / (adam_us * 1e-6) / 1e9 / peak}

    # ---- the same SpMM kernel where HBM, not L2, is the binding roofline: C5-style graph (configs[4] scaled 5x down:
    #      2.2 M nodes, 40 M edges, d = 128 -> 1.1 GB table >> 126 MB L2, every neighbour row is fetched from HBM) ----
    hbm_regime = None
    if not args.no_hbm_regime:
        from kgat_b200 import synthetic
        from kgat_b200.graph import AttentiveGraph

        n5, d5 = 2_200_000, 128
        h5, _, t5 = synthetic.make_edges_only(n5, 40_000_000, 64)
        deg5 = np.bincount(h5, minlength=n5).astype(np.float32)
        g5 = AttentiveGraph.from_coo(torch.from_numpy(h5.astype(np.int64)).to(dev), torch.from_numpy(t5.astype(np.int64)).to(dev),
                                     torch.from_numpy((1.0 / deg5[h5]).astype(np.float32)).to(dev), n5)
        x5 = torch.randn(n5, d5, device=dev)
        y5 = torch.empty_like(x5)
        for _ in range(3):
            g5.matmul(x5, out=y5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            g5.matmul(x5, out=y5)
        e1.record()
        torch.cuda.synchronize()
        t5s = e0.elapsed_time(e1) / 1e3 / reps
        b_gather = 8.0 * g5.nnz + 16.0 * g5.plan.n_tasks + 4.0 * g5.nnz * d5 + 4.0 * n5 * d5  # edges + one 512 B row per edge + output
        hbm_regime = {"kernel": "spmm_d128", "nodes": n5, "nnz": g5.nnz, "d": d5, "table_bytes": 4.0 * n5 * d5, "ms": t5s * 1e3,
                      "algorithmic_bytes": b_gather, "achieved": b_gather / t5s / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": b_gather / t5s / 1e9 / peak, "edges_per_s": g5.nnz / t5s,
                      "model": "B_gather (SURVEY.md 8d: the table is 9x the L2, so one neighbour-row read per edge is compulsory HBM traffic)"}
        del g5, x5, y5

    # ---- CPU baseline (oracle port, bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline:
        c = cpu_reference_sample(g, init_state, data)
        cpu = {"value": c["epoch_s_extrapolated"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "cf_step_s": c["cf_step_s"], "kg_step_s": c["kg_step_s"], "refresh_s": c["refresh_s_extrapolated"]}
        try:  # the same port on the GPU itself (stock ATen kernels): reported next to the CPU figure
            cpu["same_port_on_gpu"] = aten_gpu_sample(g, init_state, data, dev)
        except Exception as e:  # noqa: BLE001 - a baseline must never take the bench line down
            cpu["same_port_on_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    n_layers = 3
    edges_per_epoch = graph.nnz * n_layers * 2 * data.n_cf
    line = {
        "metric": METRIC, "value": epoch_s, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": epoch_s * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(g, data, 1),
        "propagation_edges_per_s": edges_per_epoch / epoch_s,
        "cf_loss": losses[0], "kg_loss": losses[1],
        "phases": {"cf_phase_s": phases["cf"] / 1e3, "kg_phase_s": phases["kg"] / 1e3, "refresh_s": phases["refresh"] / 1e3,
                   "cf_step_us": 1e3 * phases["cf"] / max(phases["n_cf"], 1), "kg_step_us": 1e3 * phases["kg"] / max(phases["n_kg"], 1)},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        "roofline": roofline, "roofline_hbm_regime": hbm_regime, "roofline_adam": roofline_adam, "cpu_baseline": cpu,
        "kernels": {k: {kk: round(vv, 3) if isinstance(vv, float) else vv for kk, vv in v.items()} for k, v in sorted(per_epoch_ms.items(), key=lambda kv: -kv[1]["epoch_ms"])},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
