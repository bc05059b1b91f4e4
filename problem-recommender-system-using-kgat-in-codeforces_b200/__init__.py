"""kgat_b200 -- B200-native KGAT hot path (attentive propagation over the collaborative knowledge
graph, BPR / TransR losses, attention refresh, predict) behind the reference's model API.

The package directory is named after the reference repository and contains hyphens; import it as
``kgat_b200`` (the alias module at the repo root) or via ``importlib.import_module``.

Layout
    csrc/                  hand-written sm_100a CUDA kernels + the C ABI (include/kgat_b200.h)
    lib/libkgat_b200.so    built by ``__graft_entry__.build()`` / ``make -C csrc`` (not in git)
    _lib.py, ops.py        ctypes binding and tensor-level wrappers (no fallback path)
    graph.py               CSR / CSC containers, SpMM plan, refresh edge index
    functions.py           autograd.Functions over the fused forward / backward kernels
    model.py               KGAT, KGATArgs, KGATMode  (drop-in for src.model.KGAT.model)
    aggregator.py, multi_head_attention.py, optim.py
    ckg.py, synthetic.py   host-side CKG assembly and the seeded synthetic graphs of BASELINE.json
    engine.py, trainer.py  CUDA-graph training engine, epoch driver
    metrics.py             device-side evaluate loop (top-K + precision / recall / nDCG)
    sharding.py            row-sharded multi-GPU propagation (NCCL all-gather per layer)
"""

from . import _lib, ckg, synthetic  # noqa: F401
from ._lib import KgatLibraryError  # noqa: F401
from .aggregator import Aggregator, AggregatorArgs  # noqa: F401
from .model import KGAT, KGATArgs, KGATMode  # noqa: F401
from .multi_head_attention import MultiHeadAttention  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from . import engine, functions, graph, metrics, model, ops, optim, sampler, sharding, trainer  # noqa: F401,E402

__all__ = [
    "KGAT", "KGATArgs", "KGATMode", "Aggregator", "AggregatorArgs", "MultiHeadAttention", "FusedAdam",
    "KgatLibraryError",
]
