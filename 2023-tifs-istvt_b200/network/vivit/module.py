"""Parameter holders of the decomposed spatial-temporal transformer blocks.

Same class and attribute names as the reference's `network/vivit/module.py` (`PreNorm` :15-21, `FeedForward`
:23-34, `SpatialOnlyAttention` :66-93, `TemporalResidualAttention` :174-208) so `state_dict` keys match.
The fused forward is driven by `engine.ISTVTEngine`; these modules only own the weights.

The ablation blocks `Attention` (:36-64) and `TemporalOnlyAttention` (:145-172) additionally have a stand-alone
inference `forward(x)` on CUDA token tensors, scheduled by `ablation.py`.
`LocalSpatialAttention` (:96-143) is not mirrored: its `x[:, :, 1:, :].squeeze()` / `cls_token = x[:, :, 0, :]` pair takes
the first PATCH as the class token and nothing in the reference constructs it.
"""
from __future__ import annotations

from torch import nn


class PreNorm(nn.Module):
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(nn.Module):
    def __init__(self, dim: int, hidden_dim: int, dropout: float = 0.0):
        super().__init__()
        # indices 0 and 3 carry the weights, like the reference's Sequential(Linear, GELU, Dropout, Linear, Dropout)
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class TemporalResidualAttention(nn.Module):
    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qk = nn.Linear(dim, inner * 2, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))


class SpatialOnlyAttention(nn.Module):
    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    """Joint self-attention over all tokens of a sequence (reference module.py:36-64).  `precision` ("bf16" | "fp32")
    is an addition with a preserving default."""

    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        if dim_head != 64:
            raise ValueError("the B200 attention kernels are built for dim_head = 64")
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        # project_out is always true for heads > 1 (module.py:40)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.precision = "bf16"

    def forward(self, x):
        from ...ablation import attention_forward
        return attention_forward(self, x, self.precision)


class TemporalOnlyAttention(nn.Module):
    """Attention across the frames of a clip at each of the 362 token positions (reference module.py:145-172)."""

    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        if dim_head != 64:
            raise ValueError("the B200 attention kernels are built for dim_head = 64")
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.precision = "bf16"

    def forward(self, x):
        from ...ablation import temporal_only_attention_forward
        return temporal_only_attention_forward(self, x, self.precision)
