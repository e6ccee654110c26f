"""TEST INFRASTRUCTURE — CPU oracle of the ablation transformers.

Plain-op fp32 restatement, written against a state_dict (reference key names), of
  * `Attention.forward`               network/vivit/module.py:52-63
  * `TemporalOnlyAttention.forward`   network/vivit/module.py:160-172
  * `Transformer.forward`             network/vivit/vivit.py:21-25
  * `ViViT.forward`                   network/vivit/vivit.py:60-81
  * `VanillaTr.forward`               network/vivit/vivit.py:179-191
Only tests/ may import it; the product package never does.

Pinning: oracle/make_golden_ablation.py ran the UNMODIFIED reference classes (through oracle/reference_shim.py) on
seeded inputs / weights and stored logits and fingerprints in tests/golden/ablation_golden.pt; tests/test_oracle.py
checks this file against that fixture and, where /root/reference exists, against the live reference classes.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .istvt_oracle import SD, TOKENS_PER_FRAME, _ln, entry_flow, feed_forward


def joint_attention(sd: SD, prefix: str, xn: torch.Tensor, heads: int = 8) -> torch.Tensor:
    """Attention.forward, module.py:52-63: softmax over ALL tokens of the sequence."""
    b, n, _ = xn.shape
    qkv = F.linear(xn, sd[prefix + ".to_qkv.weight"]).chunk(3, dim=-1)                    # :54
    split = lambda t: t.reshape(b, n, heads, -1).permute(0, 2, 1, 3)                      # 'b n (h d) -> b h n d' (:55)
    q, k, v = map(split, qkv)
    dots = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)                   # :57, scale :43
    attn = dots.softmax(dim=-1)                                                           # :59
    out = torch.matmul(attn, v)                                                           # :61
    out = out.permute(0, 2, 1, 3).reshape(b, n, -1)                                       # 'b h n d -> b n (h d)' (:62)
    return F.linear(out, sd[prefix + ".to_out.0.weight"], sd[prefix + ".to_out.0.bias"])  # :63


def temporal_only_attention(sd: SD, prefix: str, xn: torch.Tensor, heads: int = 8) -> torch.Tensor:
    """TemporalOnlyAttention.forward, module.py:160-172: per (clip, head, token position) attention across frames."""
    b, n, _ = xn.shape
    p = TOKENS_PER_FRAME
    f = n // p
    qkv = F.linear(xn, sd[prefix + ".to_qkv.weight"]).chunk(3, dim=-1)                    # :162
    split = lambda t: t.reshape(b, f, p, heads, -1).permute(0, 3, 2, 1, 4)               # 'b (t hw) (h d) -> b h hw t d'
    q, k, v = map(split, qkv)                                                             # :163
    dots = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)                   # :165
    attn = dots.softmax(dim=-1)                                                           # :167
    out = torch.matmul(attn, v)                                                           # :169
    out = out.permute(0, 3, 2, 1, 4).reshape(b, n, -1)                                    # 'b h hw t d -> b (t hw) (h d)'
    return F.linear(out, sd[prefix + ".to_out.0.weight"], sd[prefix + ".to_out.0.bias"])  # :171


def _depth(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.layers.{n}.0.norm.weight" in sd:
        n += 1
    return n


def plain_transformer(sd: SD, prefix: str, x: torch.Tensor, taps: Optional[dict] = None, tag: str = "") -> torch.Tensor:
    """Transformer.forward, vivit.py:21-25."""
    for layer in range(_depth(sd, prefix)):
        lp = f"{prefix}.layers.{layer}"
        x = joint_attention(sd, f"{lp}.0.fn", _ln(sd, f"{lp}.0.norm", x)) + x             # vivit.py:23
        x = feed_forward(sd, f"{lp}.1.fn", _ln(sd, f"{lp}.1.norm", x)) + x                # vivit.py:24
        if taps is not None:
            taps[f"{tag}layer{layer}"] = x
    return _ln(sd, f"{prefix}.norm", x)                                                   # vivit.py:25


def _p(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


def vivit_forward(sd: SD, feats: torch.Tensor, prefix: str = "", taps: Optional[dict] = None,
                  pool: str = "cls") -> torch.Tensor:
    """ViViT.forward, vivit.py:60-81.  feats [b, t, C, h, w] -> logits [b, num_classes]."""
    b, t, c, h, w = feats.shape
    x = feats.permute(0, 1, 3, 4, 2).reshape(b, t, h * w, c)                              # Rearrange :41 (patch size 1), :61
    n = h * w
    space = sd[_p(prefix, "space_token")].reshape(1, 1, 1, c).expand(b, t, 1, c)          # :64
    x = torch.cat((space, x), dim=2)                                                      # :65
    x = x + sd[_p(prefix, "pos_embedding")][:, :, : n + 1]                                # :66
    x = x.reshape(b * t, n + 1, c)                                                        # :69
    x = plain_transformer(sd, _p(prefix, "space_transformer"), x, taps, "space.")         # :70
    if taps is not None:
        taps["space_out"] = x
    x = x[:, 0].reshape(b, t, c)                                                          # :71
    temporal = sd[_p(prefix, "temporal_token")].reshape(1, 1, c).expand(b, 1, c)          # :73
    x = torch.cat((temporal, x), dim=1)                                                   # :74
    x = plain_transformer(sd, _p(prefix, "temporal_transformer"), x, taps, "temporal.")   # :76
    if taps is not None:
        taps["temporal_out"] = x
    x = x.mean(dim=1) if pool == "mean" else x[:, 0]                                        # :79
    return F.linear(_ln(sd, _p(prefix, "mlp_head.0"), x), sd[_p(prefix, "mlp_head.1.weight")],
                    sd[_p(prefix, "mlp_head.1.bias")])                                    # :81


def vanilla_forward(sd: SD, feats: torch.Tensor, prefix: str = "", taps: Optional[dict] = None) -> torch.Tensor:
    """VanillaTr.forward, vivit.py:179-191.  feats [b, t, C, h, w] -> logits [b, num_classes]."""
    b, t, c, h, w = feats.shape
    x = feats.permute(0, 1, 3, 4, 2).reshape(b, t, h * w, c)                              # :162
    x = F.linear(x, sd[_p(prefix, "to_patch_embedding.1.weight")], sd[_p(prefix, "to_patch_embedding.1.bias")])  # :163
    x = x.reshape(b, t * h * w, -1)                                                       # :164
    cls = sd[_p(prefix, "cls_token")].expand(b, 1, x.shape[-1])                           # :183
    x = torch.cat((cls, x), dim=1)                                                        # :184
    x = x + sd[_p(prefix, "pos_embedding")]                                               # :185
    x = plain_transformer(sd, _p(prefix, "transformer"), x, taps, "")                     # :187
    if taps is not None:
        taps["transformer_out"] = x
    x = x[:, 0]                                                                           # :189
    return F.linear(_ln(sd, _p(prefix, "mlp_head.0"), x), sd[_p(prefix, "mlp_head.1.weight")],
                    sd[_p(prefix, "mlp_head.1.bias")])                                    # :191


FORWARDS = {"vivit": vivit_forward, "vanilla": vanilla_forward}


def clip_forward(sd: SD, clips: torch.Tensor, variant: str) -> torch.Tensor:
    """XceptionVidTr.forward, vivit.py:202-208, with `vit` = ViViT / VanillaTr: clips [B, T, 3, H, W] -> [B, 1]."""
    b, t = clips.shape[:2]
    feats = entry_flow(sd, clips.reshape(b * t, *clips.shape[2:]))
    feats = feats.reshape(b, t, *feats.shape[1:])
    return FORWARDS[variant](sd, feats, "vit")


def sensitise_ablation_(sd: SD, seed: int = 2468) -> SD:
    """In place: randomise every LayerNorm gamma / beta (identity at default init, which would hide LN bugs) and halve the
    class tokens / positional embeddings.  Deterministic in (seed, sorted key order)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda shape: torch.rand(shape, generator=g)
    for k in sorted(sd.keys()):
        v = sd[k]
        if ".norm." in k or "mlp_head.0" in k:
            if k.endswith("weight"):
                v.copy_(0.75 + 0.5 * rnd(v.shape))
            elif k.endswith("bias"):
                v.copy_((rnd(v.shape) - 0.5) * 0.2)
        elif k.endswith("pos_embedding") or k.endswith("_token"):
            v.mul_(0.5)
    return sd


def make_features(b: int, t: int, seed: int = 31) -> torch.Tensor:
    """Seeded stand-in for block-3 feature maps [b, t, 728, 19, 19] (post-residual, so signed, unit scale)."""
    g = torch.Generator().manual_seed(seed + 100 * b + t)
    return torch.randn(b, t, 728, 19, 19, generator=g) * 0.8


def make_tokens(b: int, n: int, seed: int = 57) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + n)
    return torch.randn(b, n, 728, generator=g)
