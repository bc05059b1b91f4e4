"""Bi-interaction aggregator layer -- same surface as the reference ``Aggregator``
(src/model/KGAT/aggregator.py:8-65): ``AggregatorArgs(input_dim, output_dim, dropout)``,
sub-modules ``message_dropout``, ``activation``, ``linear1``, ``linear2`` (so checkpoints keep the
keys ``_aggregator_layers.{l}.linear{1,2}.{weight,bias}``) and ``forward(ego, attentive_matrix)``.

The module only *holds* the parameters; the arithmetic runs in the fused SpMM + bi-interaction
kernels (K1 + K2).  Construction order of the sub-modules mirrors the reference so that, under the
same ``torch.manual_seed``, the initial weights are identical.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from .functions import DropoutSpec, PropagateFunction
from .graph import AttentiveGraph


@dataclass
class AggregatorArgs:
    input_dim: int
    output_dim: int
    dropout: float


class Aggregator(nn.Module):
    def __init__(self, args: AggregatorArgs) -> None:
        super().__init__()
        self._input_dim = args.input_dim
        self._output_dim = args.output_dim
        self.message_dropout = nn.Dropout(p=args.dropout)
        self.activation = nn.LeakyReLU()
        self.linear1 = nn.Linear(in_features=self._input_dim, out_features=self._output_dim)
        self.linear2 = nn.Linear(in_features=self._input_dim, out_features=self._output_dim)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)

    def kernel_params(self):
        return (self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)

    def forward(self, ego_embeddings: torch.Tensor, attentive_matrix) -> torch.Tensor:
        """Stand-alone layer call (aggregator.py:37-65).  ``attentive_matrix`` may be a sparse COO
        tensor (coalesced on the fly) or an ``AttentiveGraph``."""
        graph = attentive_matrix if isinstance(attentive_matrix, AttentiveGraph) else AttentiveGraph.from_sparse_coo(attentive_matrix)
        p = self.message_dropout.p if self.training else 0.0
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p > 0 else 0
        (out,) = PropagateFunction.apply(graph, DropoutSpec(ps=[p], seed=seed), ego_embeddings.contiguous(), *self.kernel_params())
        return out
