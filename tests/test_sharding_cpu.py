"""Multi-GPU host logic on the CPU: the cyclic partition index math and -- with world_size 2 over
gloo and plain-torch local ops injected in place of the CUDA kernels -- the exchange pattern of the
row-sharded CF step (per-layer all-gather forward, all-gather of the side gradient backward,
all-reduce of the dense-parameter gradients) against the unsharded oracle."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_balanced_ranges_cover_the_nodes_and_balance_the_cost(world):
    """Host logic of the range-sharded engine (sharded_pruned.balanced_ranges): contiguous, 32-aligned node ranges that cover
    [0, N) exactly once and carry about 1/world of the cost model each, on a degree profile shaped like the CKG's (many short
    user rows first, then heavy item rows, then entity rows)."""
    from kgat_b200.sharded_pruned import balanced_ranges

    rng = np.random.default_rng(world)
    lens = np.concatenate([rng.integers(1, 30, 7000), rng.integers(50, 3000, 2500), rng.integers(1, 400, 9000)])
    t_lens = rng.permutation(lens)
    row_ptr, t_ptr = np.concatenate([[0], np.cumsum(lens)]), np.concatenate([[0], np.cumsum(t_lens)])
    n = lens.shape[0]
    r = balanced_ranges(row_ptr, t_ptr, world)
    assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
    for (lo, hi), (lo2, _) in zip(r, r[1:]):
        assert hi == lo2
    for lo, hi in r:
        assert lo % 32 == 0 and hi > lo
    row_cost = 48.0 + 64.0 * (world - 1)
    cost = lens + t_lens + row_cost
    shares = np.array([cost[lo:hi].sum() for lo, hi in r]) / cost.sum()
    assert abs(shares.sum() - 1.0) < 1e-12
    # a cut moves by at most 16 rows from the ideal one (32-alignment): the heaviest rows bound the imbalance
    slack = 2 * 32 * cost.max() / cost.sum()
    assert np.abs(shares - 1.0 / world).max() <= slack + 1e-12, (shares, slack)
    # degenerate input: fewer 32-row blocks than ranks still yields valid, ordered ranges
    tiny = balanced_ranges(np.arange(0, 41), np.arange(0, 41), 2)
    assert tiny[0][0] == 0 and tiny[-1][1] == 40 and tiny[0][1] == tiny[1][0]


def test_cyclic_partition_index_math():
    from kgat_b200.sharding import CyclicPartition, shard_csr

    n, world = 23, 4
    parts = [CyclicPartition(n, world, r) for r in range(world)]
    assert parts[0].max_rows == 6 and parts[0].padded == 24
    assert sum(p.count() for p in parts) == n
    ids = np.arange(n)
    pad = parts[0].to_padded(ids)
    assert len(set(pad.tolist())) == n and pad.max() < 24
    for p in parts:
        rows = p.local_rows()
        assert (rows % world == p.rank).all()
        sl = p.slice()
        np.testing.assert_array_equal(p.to_padded(rows), np.arange(sl.start, sl.stop))
    t = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    assert torch.equal(parts[2].gather_rows(parts[2].scatter_rows(t)), t)
    # local CSR: every non-zero lands on exactly one rank, columns remapped to the padded layout
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 6, n)
    rp = np.concatenate([[0], np.cumsum(lens)])
    ci = rng.integers(0, n, rp[-1])
    seen = np.zeros(rp[-1], int)
    for p in parts:
        lp, lc, slots = shard_csr(rp, ci, p)
        seen[slots] += 1
        np.testing.assert_array_equal(lc, p.to_padded(ci[slots]))
        np.testing.assert_array_equal(np.diff(lp), lens[p.local_rows()])
    assert (seen == 1).all()


def test_cyclic_partition_properties_random_shapes():
    """Property test over random (n, world, chunking): the padded cyclic layout is a bijection onto distinct rows, rank
    slices tile it, and the row-chunked local CSRs partition the non-zeros of a rank exactly."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from kgat_b200.sharding import CyclicPartition, shard_csr

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 200), world=st.integers(1, 9), n_chunks=st.integers(1, 5), seed=st.integers(0, 10**6))
    def check(n, world, n_chunks, seed):
        rng = np.random.default_rng(seed)
        parts = [CyclicPartition(n, world, r) for r in range(world)]
        pad = parts[0].to_padded(np.arange(n))
        assert len(set(pad.tolist())) == n and pad.min() >= 0 and pad.max() < parts[0].padded
        assert sum(p.count() for p in parts) == n and all(p.count() <= p.max_rows for p in parts)
        lens = rng.integers(0, 5, n)
        rp = np.concatenate([[0], np.cumsum(lens)])
        ci = rng.integers(0, n, rp[-1])
        seen = np.zeros(rp[-1], int)
        for p in parts:
            sl, fs = p.slice(), p.full_slice()
            assert fs.start == p.rank * p.max_rows and fs.stop - fs.start == p.max_rows and sl.stop - sl.start == p.count()
            cuts = [p.count() * c // n_chunks for c in range(n_chunks + 1)]
            whole = shard_csr(rp, ci, p)
            got_slots, got_cols = [], []
            for c in range(n_chunks):
                lp, lc, slots = shard_csr(rp, ci, p, (cuts[c], cuts[c + 1]))
                assert lp[0] == 0 and len(lp) == cuts[c + 1] - cuts[c] + 1 and lp[-1] == len(lc) == len(slots)
                np.testing.assert_array_equal(np.diff(lp), lens[p.local_rows()[cuts[c] : cuts[c + 1]]])
                got_slots.append(slots)
                got_cols.append(lc)
                seen[slots] += 1
            np.testing.assert_array_equal(np.concatenate(got_slots) if got_slots else [], whole[2])
            np.testing.assert_array_equal(np.concatenate(got_cols) if got_cols else [], whole[1])
        assert (seen == 1).all()
        t = torch.from_numpy(rng.standard_normal((n, 2)).astype(np.float32))
        assert torch.equal(parts[-1].gather_rows(parts[-1].scatter_rows(t)), t)

    check()


class TorchOps:
    """LocalOps with plain torch CPU ops (test double for sharding.KernelOps)."""

    def spmm(self, lg, x_full, out, addend=None):
        y = torch.sparse.mm(lg["csr"], x_full)
        out.copy_(y if addend is None else y + addend)

    def biagg_forward(self, e, s, layer, out, p, seed, offset, seed_dev=None):
        assert p == 0.0
        w1, b1, w2, b2 = layer
        x = F.leaky_relu(F.linear(e + s, w1, b1), 0.01) + F.leaky_relu(F.linear(e * s, w2, b2), 0.01)
        out.copy_(F.normalize(x, dim=1, eps=1e-12))
        return None, None

    def biagg_backward(self, g_out, out, inv, flags, e, s, layer, p, g_s, g_e, accumulate_into=None):
        leaves = [t.detach().clone().requires_grad_(True) for t in (e, s, *layer)]
        ee, ss, w1, b1, w2, b2 = leaves
        x = F.leaky_relu(F.linear(ee + ss, w1, b1), 0.01) + F.leaky_relu(F.linear(ee * ss, w2, b2), 0.01)
        F.normalize(x, dim=1, eps=1e-12).backward(g_out)
        g_e.copy_(ee.grad)
        g_s.copy_(ss.grad)
        grads = [w1.grad, b1.grad, w2.grad, b2.grad]
        if accumulate_into is not None:  # later row chunks of a layer add to the first chunk's parameter gradients
            for acc, g in zip(accumulate_into, grads):
                acc.add_(g)
            return accumulate_into
        return grads

    def bpr_forward(self, tables, u, p, n, reg, loss, scratch):
        from oracle import kgat_oracle as O

        loss.copy_(O.bpr_loss_from_table(torch.cat(tables, dim=1), u, p, n, reg).reshape(1))

    def bpr_backward(self, tables, grads, u, p, n, reg, scratch, g_loss):
        from oracle import kgat_oracle as O

        leaves = [t.detach().clone().requires_grad_(True) for t in tables]
        (O.bpr_loss_from_table(torch.cat(leaves, dim=1), u, p, n, reg) * g_loss[0]).backward()
        for g, leaf in zip(grads, leaves):
            if g is not None:
                g.add_(leaf.grad)


def _worker(rank, world, port, ok, n_chunks=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import Golden
        from kgat_b200.sharding import CyclicPartition, ShardedPropagation, shard_csr
        from oracle import kgat_oracle as O

        g = Golden("model_tiny.npz")
        n = g.node_num
        params = g.params()
        att = g.att_coo().coalesce()
        part = CyclicPartition(n, world, rank)
        crow = torch._convert_indices_from_coo_to_csr(att.indices()[0], n).numpy()
        col, val = att.indices()[1].numpy(), att.values()
        att_t = att.t().coalesce()
        trow = torch._convert_indices_from_coo_to_csr(att_t.indices()[0], n).numpy()
        tcol, tval = att_t.indices()[1].numpy(), att_t.values()

        def local(rp, ci, vals, rng=None):
            lp, lc, slots = shard_csr(rp, ci, part, rng)
            rows = part.count() if rng is None else rng[1] - rng[0]
            csr = torch.sparse_csr_tensor(torch.from_numpy(lp).long(), torch.from_numpy(lc).long(), vals[torch.from_numpy(slots).long()],
                                          size=(rows, part.padded))
            return {"csr": csr}

        dims = [64, 64, 32, 16]
        # the forward SpMM / bi-interaction optionally run over consecutive chunks of the local rows (the pipeline the
        # peer exchange uses to overlap transfers); the result must not depend on the chunking
        cuts = [part.count() * c // n_chunks for c in range(n_chunks + 1)]
        bounds = [(cuts[c], cuts[c + 1]) for c in range(n_chunks)]
        a_chunks = [local(crow, col, val, b) for b in bounds]
        prop = ShardedPropagation(part, a_chunks, local(trow, tcol, tval), TorchOps(), dims, "cpu", chunk_bounds=bounds)
        layers = [tuple(params[f"_aggregator_layers.{l}.linear{k}.{w}"] for k in (1, 2) for w in ("weight", "bias")) for l in range(3)]
        e0 = params["_user_entity_embedding.weight"]
        # only the own slice is filled: the forward all-gathers the rest
        prop.tables[0][part.slice()] = e0[torch.from_numpy(part.local_rows())]
        u, p, q = (torch.from_numpy(g[k]) for k in ("cf_users", "cf_pos", "cf_neg"))
        loss = torch.zeros(1)
        prop.forward(layers, [0.0] * 3, 0, part.to_padded(u), part.to_padded(p), part.to_padded(q), 1e-5, loss, None)
        g_e0_loc, pgrads = prop.backward(layers, torch.ones(1))

        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        ref = O.cf_loss(leaves, att, u, p, q)
        ref.backward()
        assert abs(float(loss) - float(ref)) < 1e-6, (float(loss), float(ref))
        assert abs(float(loss) - float(g["cf_loss_eval"])) < 1e-6  # and against the reference golden
        ref_g = leaves["_user_entity_embedding.weight"].grad[torch.from_numpy(part.local_rows())]
        assert float((g_e0_loc - ref_g).abs().max()) < 1e-7 * max(1.0, float(ref_g.abs().max()) * 1e3)
        for l in range(3):
            for j, (k, w) in enumerate(((1, "weight"), (1, "bias"), (2, "weight"), (2, "bias"))):
                rg = leaves[f"_aggregator_layers.{l}.linear{k}.{w}"].grad
                assert float((pgrads[l][j] - rg).abs().max()) <= 2e-5 * float(rg.abs().max()) + 1e-9, (l, k, w)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n_chunks", [1, 3])
def test_sharded_cf_step_world2_gloo_matches_oracle(n_chunks):
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = 29000 + (os.getpid() % 2000) + 7 * n_chunks
    mp.spawn(_worker, args=(world, port, ok, n_chunks), nprocs=world, join=True)
    assert list(ok) == [1] * world
