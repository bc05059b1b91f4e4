"""Launched under torchrun by test_gpu_parity.py::test_sharded_engine_two_gpus (and by hand):
the row-sharded CF phase on WORLD_SIZE GPUs against the single-GPU engine, same seeded model, dropout off
(each rank draws its own dropout stream), for every exchange implementation.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/sharded_check.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kgat_b200 import sharding, synthetic  # noqa: E402
from kgat_b200.engine import TrainEngine  # noqa: E402
from kgat_b200.model import KGATMode  # noqa: E402
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
    n_cf = 6
    data = EpochData.sample(g, seed=3, n_cf=n_cf, n_kg=4)
    ok = True

    def fresh():
        m = build_model(g, dev, seed=11, message_dropout=[0.0, 0.0, 0.0])
        holder = TrainEngine(m, use_graphs=False).bind_resident(data.tensors())
        m(*holder.edges, mode=KGATMode.UPDATE_ATTENTION)
        return m, holder

    ref, holder = fresh()
    ref_eng = TrainEngine(ref)
    ref_eng.bind_resident(data.tensors())
    ref_losses = ref_eng.run_epoch()
    ref_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if not v.is_sparse}
    for kind in ("nccl", "peer", "peer-fwd", "peer-all"):
        for use_graphs in (False, True):
            m, holder = fresh()
            eng = sharding.ShardedEngine(m, sharding.CyclicPartition(g.node_num, world, rank), use_graphs=use_graphs, exchange=kind)
            losses = eng.run_epoch(holder)
            errs = {k: rel(v, ref_state[k]) for k, v in m.state_dict().items() if not v.is_sparse}
            worst = max(errs.values())
            # all ranks must hold identical replicas afterwards
            w = m._user_entity_embedding.weight.detach()
            chk = torch.stack([w.double().sum(), w.double().abs().sum()])
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            same = all(torch.equal(allc[0], c) for c in allc)
            good = worst < 2e-4 and abs(losses[0] - ref_losses[0]) < 1e-5 and abs(losses[1] - ref_losses[1]) < 1e-5 and same
            ok = ok and good
            if rank == 0:
                print(f"exchange={kind:9s} graphs={use_graphs!s:5s} worst rel err {worst:.2e}  cf {losses[0]:.6f} (ref {ref_losses[0]:.6f})  "
                      f"kg {losses[1]:.6f} (ref {ref_losses[1]:.6f})  replicas equal {same}  {'OK' if good else 'FAIL'}", file=sys.stderr)
            eng._cf_graph = None
            eng.single._graphs.clear()
            torch.cuda.synchronize()
            dist.barrier()
            if eng.exchange is not None:
                eng.exchange.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    sys.stderr.flush()
    os._exit(0 if int(flag.item()) == 0 else 1)


main()
