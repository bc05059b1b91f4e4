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
        """Stand-alone call with the reference's signature and output shape ``(batch, 1, kg_embedding_dim)``
        (multi_head_attention.py:35-58).  ``head_embedding`` / ``relation_embedding`` are accepted and shape-checked but
        cannot influence the result: the reference's softmax runs over a length-1 key axis.  Per-head
        dropout is live in ``train()`` mode, as in the reference.  No autograd graph is recorded: in the reference these
        weights never receive a gradient either (the refreshed matrix is assigned through ``.data``, model.py:366)."""
        from . import ops

        batch = head_embedding.size(0)
        if tail_embedding.size(0) != batch or tail_embedding.size(-1) != self._cf_embedding_dim or relation_embedding.size(-1) != self._cf_embedding_dim:
            raise ValueError("MultiHeadAttention.forward: (batch, cf_dim) heads / tails and a (cf_dim,) relation embedding are expected")
        x = tail_embedding.detach().reshape(batch, self._cf_embedding_dim).to(torch.float32).contiguous()
        p = self.dropout_p if self.training else 0.0
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p > 0 else 0
        head_bits = getattr(self, "_injected_head_bits", None)  # test hook: uint8 [batch], bit h keeps head h
        y = ops.mha_forward(x, self.kernel_params(), dropout_p=p, head_bits=head_bits, seed=seed, n_heads=self._head_num, eps=self.ln_eps)
        return y.view(batch, 1, self._kg_embedding_dim)
