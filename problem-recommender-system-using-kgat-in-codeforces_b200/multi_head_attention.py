"""Parameter holder with the reference's ``MultiHeadAttention`` layout
(src/model/KGAT/multi_head_attention.py:5-29): four ``nn.Linear`` (query / key / value / output), a
``LayerNorm`` and a ``Dropout(0.1)`` over 8 heads, xavier-initialised in the same order so seeds and
checkpoints (``_multi_head_attention.*`` keys) are interchangeable.

In the reference this module scores every CKG edge during the attention refresh.  Its softmax runs
over a length-1 key axis, so the query / key projections cancel and the output depends only on the
value path.  The arithmetic therefore lives in the fused per-pair /
per-edge kernels (``csrc/attention.cu``), which read ``_value_weight``, ``_output`` and
``_layer_norm`` from here; ``_query_weight`` / ``_key_weight`` are kept only for checkpoint
compatibility.
"""

from __future__ import annotations

import torch
from torch import nn


class MultiHeadAttention(nn.Module):
    def __init__(self, cf_embedding_dim: int, kg_embedding_dim: int, head_num: int = 8, dropout: float = 0.1) -> None:
        super().__init__()
        self._head_num = head_num
        self._cf_embedding_dim = cf_embedding_dim
        self._kg_embedding_dim = kg_embedding_dim
        self._depth = self._kg_embedding_dim // self._head_num
        self._query_weight = nn.Linear(self._cf_embedding_dim, self._kg_embedding_dim)
        self._key_weight = nn.Linear(self._cf_embedding_dim, self._kg_embedding_dim)
        self._value_weight = nn.Linear(self._cf_embedding_dim, self._kg_embedding_dim)
        self._output = nn.Linear(self._kg_embedding_dim, self._kg_embedding_dim)
        self._layer_norm = nn.LayerNorm(self._kg_embedding_dim)
        self._dropout = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self._query_weight.weight)
        nn.init.xavier_uniform_(self._key_weight.weight)
        nn.init.xavier_uniform_(self._value_weight.weight)
        nn.init.xavier_uniform_(self._output.weight)

    def kernel_params(self) -> dict:
        """Tensors the attention kernels consume (all fp32, contiguous)."""
        return {
            "Wv": self._value_weight.weight.detach(),
            "bv": self._value_weight.bias.detach(),
            "Wo": self._output.weight.detach(),
            "bo": self._output.bias.detach(),
            "gamma": self._layer_norm.weight.detach(),
            "beta": self._layer_norm.bias.detach(),
        }

    @property
    def head_num(self) -> int:
        return self._head_num

    @property
    def dropout_p(self) -> float:
        return float(self._dropout.p)

    @property
    def ln_eps(self) -> float:
        return float(self._layer_norm.eps)

    def forward(self, head_embedding: torch.Tensor, relation_embedding: torch.Tensor, tail_embedding: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "kgat_b200 fuses this module into the attention-refresh kernels (KGATMode.UPDATE_ATTENTION); "
            "it has no stand-alone forward and no PyTorch fallback."
        )
