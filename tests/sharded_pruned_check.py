"""Launched under torchrun by test_gpu_pruning.py::test_range_sharded_engine_multi_gpu (and by hand): the range-sharded pruned
epoch (sharded_pruned.RangeShardedEngine) on WORLD_SIZE GPUs against the single-GPU engine -- same seeded model, message dropout
off -- eager and captured, plus identical replicas on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/sharded_pruned_check.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import synthetic  # noqa: E402
from kgat_b200.engine import TrainEngine  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
from kgat_b200.sharded_pruned import RangeShardedEngine  # noqa: E402
from kgat_b200.trainer import EpochData, build_model  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    os.dup2(2, 1)  # NCCL banner
    shape = os.environ.get("KGAT_CHECK_SHAPE", "small")
    g = synthetic.make_ckg(shape, seed=5)
    data = EpochData.sample(g, seed=3, n_cf=6, n_kg=4)
    ok = True

    def fresh():
        m = build_model(g, dev, seed=11, message_dropout=[0.0, 0.0, 0.0])
        m._multi_head_attention._dropout.p = 0.0
        holder = TrainEngine(m, use_graphs=False).bind_resident(data.tensors())
        m(*holder.edges, mode=KGATMode.UPDATE_ATTENTION)
        return m

    ref = fresh()
    ref_eng = TrainEngine(ref)
    ref_eng.bind_resident(data.tensors())
    ref_losses = [ref_eng.run_epoch()[:2] for _ in range(2)]
    ref_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if not v.is_sparse}
    for use_graphs in (False, True):
        m = fresh()
        eng = RangeShardedEngine(m, world, rank, use_graphs=use_graphs)
        eng.bind_resident(data.tensors())
        losses = [eng.run_epoch(epoch_seed=i) for i in range(2)]  # two epochs: the second runs on the refreshed graph
        errs = {k: rel(v, ref_state[k]) for k, v in m.state_dict().items() if not v.is_sparse}
        worst = max(errs.values())
        chk = eng.replica_checksum
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        dl = max(abs(a[0] - b[0]) + abs(a[1] - b[1]) for a, b in zip(losses, ref_losses))
        good = worst < 2e-4 and same and dl < 2e-5
        ok = ok and good
        if rank == 0:
            print(f"world={world} graphs={use_graphs!s:5s} worst rel err {worst:.2e} loss diff {dl:.2e} replicas identical {same} ranges {eng.bounds} -> {'ok' if good else 'FAIL'}",
                  file=sys.stderr)
        eng.close()
        del eng, m
        torch.cuda.empty_cache()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.barrier()
    sys.stderr.flush()
    os._exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
