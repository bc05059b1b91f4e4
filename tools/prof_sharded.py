#!/usr/bin/env python
"""Where a row-sharded CF step spends its time: eager (un-captured) sharded CF steps at the C3 shape with CUDA
events around every collective and every kernel family, plus the captured-graph step time for comparison.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/prof_sharded.py
"""
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import sharding, synthetic  # noqa: E402
from kgat_b200.engine import TrainEngine  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
from kgat_b200.trainer import EpochData, build_model  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
sys.stdout.flush()
os.dup2(2, 1)

g = synthetic.make_ckg("amazon-book", with_dicts=True)
data = EpochData.sample(g, n_cf=64, n_kg=8)
model = build_model(g, dev)
part = sharding.CyclicPartition(g.node_num, world, rank)
holder = TrainEngine(model, use_graphs=False).bind_resident(data.tensors())
model(*holder.edges, mode=KGATMode.UPDATE_ATTENTION)

events = defaultdict(list)


def timed(name, fn):
    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        events[name].append((e0, e1))
        return out

    return wrapper


def report(tag, steps):
    torch.cuda.synchronize()
    rows = sorted(((sum(a.elapsed_time(b) for a, b in v) * 1e3 / steps, len(v) // steps, k) for k, v in events.items()), reverse=True)
    if rank == 0:
        print(f"--- {tag}: per step, rank 0 of {world}", file=sys.stderr)
        for us, n, k in rows:
            print(f"{k:28s} {us:9.1f} us  ({n} calls)", file=sys.stderr)
        print(f"{'sum':28s} {sum(r[0] for r in rows):9.1f} us", file=sys.stderr)
    events.clear()


# graphed step time first
eng = sharding.ShardedEngine(model, part)
for _ in range(2):
    eng.run_epoch(holder, n_kg=0, refresh=False)
torch.cuda.synchronize()
dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
eng.run_epoch(holder, n_kg=0, refresh=False)
t1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"[{eng.exchange_kind}] graphed sharded CF step: {t0.elapsed_time(t1) * 1e3 / data.n_cf:.1f} us (incl. per-epoch gather/scatter)",
          file=sys.stderr)
ref = eng.prop.tables[0].clone()

eng = sharding.ShardedEngine(model, part, use_graphs=False)
ex = eng.prop.ex
ex.gather = timed("exchange rows", ex.gather)
ex.all_reduce_flat = timed("all_reduce grads", ex.all_reduce_flat)
k = eng.kops
for name in ("spmm", "biagg_forward", "biagg_backward", "bpr_forward", "bpr_backward"):
    setattr(k, name, timed(name, getattr(k, name)))
eng.ops.adam_apply = timed("adam", eng.ops.adam_apply)
eng.cf_step = timed("cf_step_total", eng.cf_step)
eng.run_epoch(holder, n_cf=8, n_kg=0, refresh=False)
events.clear()
dist.barrier()
eng.run_epoch(holder, n_cf=32, n_kg=0, refresh=False)
report("eager", 32)
# the two engines ran different numbers of steps from the same start; cross-rank consistency check instead:
chk = torch.stack([t.double().sum() for t in eng.prop.tables])
all_chk = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(all_chk, chk)
if rank == 0:
    same = [bool(all(c[i] == all_chk[0][i] for c in all_chk)) for i in range(chk.numel())]
    print("layer-table checksums equal across ranks (table 0 differs by design: own rows are one Adam step ahead):", same,
          [f"{v:.6e}" for v in chk.tolist()], file=sys.stderr)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
