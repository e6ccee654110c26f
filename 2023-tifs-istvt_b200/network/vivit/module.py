"""Parameter holders of the decomposed spatial-temporal transformer blocks.

Same class and attribute names as the reference's `network/vivit/module.py` (`PreNorm` :15-21, `FeedForward`
:23-34, `SpatialOnlyAttention` :66-93, `TemporalResidualAttention` :174-208) so `state_dict` keys match.
The fused forward is driven by `engine.ISTVTEngine`; these modules only own the weights.
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
