"""Forward engine of the ablation transformers: weight packing + kernel schedule of

  * `Attention`              network/vivit/module.py:36-64   joint self-attention over all tokens of a sequence
  * `TemporalOnlyAttention`  network/vivit/module.py:145-172 attention across frames at each token position
  * `Transformer`            network/vivit/vivit.py:10-25    depth x (PreNorm Attention + residual, PreNorm MLP + residual)
  * `ViViT`                  network/vivit/vivit.py:29-81    factorised encoder: per-frame transformer, then per-clip
  * `VanillaTr`              network/vivit/vivit.py:150-191  one joint transformer over the T*361+1 tokens of a clip

on the kernels of the ISTVT path: the same tcgen05 GEMM (+ bias / GELU / in-place residual epilogues), LayerNorm, head
and temporal-attention kernels; sequences of up to 384 tokens use the spatial-attention kernel (all keys resident in
TMEM), longer ones the key-streaming `istvt_attn_joint_fwd`.  The residual stream is fp32 [sequences, tokens, dim] and is
updated in place; activations are bf16 (fp32 in the validation mode).  Inference only; no torch-op or CPU fallback.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .engine import PRECISIONS, _f32, pack_entry, run_entry_flow

SPATIAL_KERNEL_MAX_TOKENS = 384      # csrc/attn_spatial.cu SA_KMAX


@dataclass
class _TLayerPack:
    ln1: Tuple[torch.Tensor, torch.Tensor]
    w_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    ln2: Tuple[torch.Tensor, torch.Tensor]
    w_1: torch.Tensor
    b_1: torch.Tensor
    w_2: torch.Tensor
    b_2: torch.Tensor


@dataclass
class _TransformerPack:
    layers: List[_TLayerPack]
    norm: Tuple[torch.Tensor, torch.Tensor]
    heads: int
    scale: float = 64 ** -0.5          # Attention.scale = dim_head ** -0.5 (module.py:41)
    fingerprint: tuple = field(default_factory=tuple)


def _fingerprint(module) -> tuple:
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def _ln(m) -> Tuple[torch.Tensor, torch.Tensor]:
    return _f32(m.weight), _f32(m.bias)


def pack_transformer(tr, dt: torch.dtype) -> _TransformerPack:
    """tr: `Transformer` parameter tree (layers[i] = [PreNorm(Attention), PreNorm(FeedForward)], vivit.py:16-19)."""
    wt = lambda m: m.weight.detach().to(dt).contiguous()
    layers = []
    heads, scale = 8, 64 ** -0.5
    for attn, ff in tr.layers:
        heads, scale = attn.fn.heads, float(attn.fn.scale)
        layers.append(_TLayerPack(
            ln1=_ln(attn.norm), w_qkv=wt(attn.fn.to_qkv), w_o=wt(attn.fn.to_out[0]), b_o=_f32(attn.fn.to_out[0].bias),
            ln2=_ln(ff.norm), w_1=wt(ff.fn.net[0]), b_1=_f32(ff.fn.net[0].bias),
            w_2=wt(ff.fn.net[3]), b_2=_f32(ff.fn.net[3].bias)))
    return _TransformerPack(layers=layers, norm=_ln(tr.norm), heads=heads, scale=scale, fingerprint=_fingerprint(tr))


def attention(qkv: torch.Tensor, sequences: int, tokens: int, heads: int, scale: float) -> torch.Tensor:
    """softmax(q k^T scale) v per (sequence, head) on the packed projection output (module.py:53-61)."""
    if tokens <= SPATIAL_KERNEL_MAX_TOKENS:
        return ops.attn_spatial(qkv, sequences, tokens, heads, scale)[0]
    return ops.attn_joint(qkv, sequences, tokens, heads, scale)


def run_transformer(tp: _TransformerPack, tok: torch.Tensor, sequences: int, tokens: int, dt: torch.dtype,
                    taps: Optional[dict] = None, tag: str = "", cls_only: bool = False) -> torch.Tensor:
    """`Transformer.forward` WITHOUT its final LayerNorm (vivit.py:22-24): tok fp32 [sequences*tokens, dim], updated
    in place.  The caller applies `norm` to the rows it reads (LayerNorm is row-wise).

    `cls_only`: the caller reads only token 0 of every sequence (vivit.py:71, :79, :189).  Everything after the last
    layer's attention is row-wise, so that layer's output projection, residual adds and MLP then run on the class rows
    alone and the function returns them as fp32 [sequences, dim] (`tok` is left one layer behind).  Attention itself and
    the layers before it need every row.  Not used when intermediates are tapped."""
    rows, dim = tok.shape
    scale = tp.scale
    prune = cls_only and taps is None and tokens > 1
    for li, lp in enumerate(tp.layers):
        xn = ops.layernorm(tok, lp.ln1[0], lp.ln1[1], dt)                               # PreNorm, module.py:21
        qkv = ops.gemm(xn, lp.w_qkv)                                                    # to_qkv, module.py:54
        a = attention(qkv, sequences, tokens, tp.heads, scale)
        if prune and li == len(tp.layers) - 1:
            inner = tp.heads * 64
            a_cls = ops.gather_rows(a, sequences, tokens * inner, 1, inner, inner)      # query 0 of every sequence
            x_cls = ops.gather_rows(tok, sequences, tokens * dim, 1, dim, dim)          # fp32 residual rows
            ops.gemm(a_cls, lp.w_o, bias=lp.b_o, residual=x_cls, out=x_cls)
            zn = ops.layernorm(x_cls, lp.ln2[0], lp.ln2[1], dt)
            hid = ops.gemm(zn, lp.w_1, bias=lp.b_1, act=ops.ACT_GELU, out_dtype=dt)
            ops.gemm(hid, lp.w_2, bias=lp.b_2, residual=x_cls, out=x_cls)
            return x_cls
        ops.gemm(a, lp.w_o, bias=lp.b_o, residual=tok, out=tok)                         # to_out + residual, vivit.py:23
        zn = ops.layernorm(tok, lp.ln2[0], lp.ln2[1], dt)
        hid = ops.gemm(zn, lp.w_1, bias=lp.b_1, act=ops.ACT_GELU, out_dtype=dt)         # module.py:33-34
        ops.gemm(hid, lp.w_2, bias=lp.b_2, residual=tok, out=tok)                       # vivit.py:24
        if taps is not None:
            taps[f"{tag}layer{li}"] = tok.clone()
        del xn, qkv, a, zn, hid
    if cls_only:
        return ops.gather_rows(tok, sequences, tokens * dim, 1, dim, dim)
    return tok


class _PackCache:
    """(device, precision) -> pack, rebuilt when a parameter's storage or version counter changes."""

    def __init__(self, builder):
        self._builder = builder
        self._packs: Dict[Tuple[str, str], object] = {}

    def get(self, module, dev: torch.device, precision: str):
        key = (str(dev), precision)
        pk = self._packs.get(key)
        if pk is None or pk.fingerprint != _fingerprint(module):
            pk = self._builder(module, PRECISIONS[precision])
            self._packs[key] = pk
        return pk


def _check_precision(precision: str) -> torch.dtype:
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
    return PRECISIONS[precision]


def _check_tokens(x: torch.Tensor, dim: int, what: str) -> None:
    if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[2] != dim:
        raise ValueError(f"{what} expects a token tensor [b, n, {dim}]")
    if not x.is_cuda:
        raise ValueError(f"{what} (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")


def _inference_only(fn):
    """Public entry points: refuse ANY training-mode call, then run without recording a graph.  In train mode the
    reference uses BatchNorm batch statistics (and updates the running ones) and applies dropout; these paths fold the
    running statistics into the convolutions and have no dropout kernel, so a `module.train()` call — with or without
    `torch.no_grad()` — would silently compute something else.  The backward kernels are built for the ISTVT model
    (DSTTr) only."""

    @functools.wraps(fn)
    def wrapper(module, *args, **kwargs):
        if module.training:
            raise NotImplementedError(f"istvt_b200: {type(module).__name__} is an inference path here: call "
                                      "model.eval() first (train mode means BatchNorm batch statistics and dropout in "
                                      "the reference; the training step is built for the ISTVT model, DSTTr, only)")
        with torch.no_grad():
            return fn(module, *args, **kwargs)

    return wrapper


# ------------------------------------------------------------------------------------------------
# standalone blocks
# ------------------------------------------------------------------------------------------------
@dataclass
class _BlockPack:
    w_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    fingerprint: tuple = field(default_factory=tuple)


def _pack_block(mod, dt: torch.dtype) -> _BlockPack:
    return _BlockPack(w_qkv=mod.to_qkv.weight.detach().to(dt).contiguous(),
                      w_o=mod.to_out[0].weight.detach().to(dt).contiguous(), b_o=_f32(mod.to_out[0].bias),
                      fingerprint=_fingerprint(mod))


def _block_pack(mod, dev: torch.device, precision: str) -> _BlockPack:
    cache = mod.__dict__.get("_pack_cache")
    if cache is None:
        cache = _PackCache(_pack_block)
        object.__setattr__(mod, "_pack_cache", cache)
    return cache.get(mod, dev, precision)


@_inference_only
def attention_forward(mod, x: torch.Tensor, precision: str = "bf16") -> torch.Tensor:
    """`Attention.forward`, module.py:52-63: x [b, n, dim] -> [b, n, dim] fp32."""
    dt = _check_precision(precision)
    _check_tokens(x, mod.to_qkv.in_features, "Attention")
    b, n, dim = x.shape
    bp = _block_pack(mod, x.device, precision)
    xa = x.reshape(b * n, dim).to(dt).contiguous()
    qkv = ops.gemm(xa, bp.w_qkv)
    a = attention(qkv, b, n, mod.heads, mod.scale)
    return ops.gemm(a, bp.w_o, bias=bp.b_o, out_dtype=torch.float32).view(b, n, dim)


@_inference_only
def temporal_only_attention_forward(mod, x: torch.Tensor, precision: str = "bf16") -> torch.Tensor:
    """`TemporalOnlyAttention.forward`, module.py:160-172: x [b, t*362, dim] -> [b, t*362, dim] fp32.  The fused to_qkv
    weight is split into its q|k rows and its v rows so that the temporal kernel reads both projections in place."""
    dt = _check_precision(precision)
    _check_tokens(x, mod.to_qkv.in_features, "TemporalOnlyAttention")
    b, n, dim = x.shape
    p_tok = 19 * 19 + 1                                                  # hard-coded in the reference, module.py:163
    if n % p_tok:
        raise ValueError(f"TemporalOnlyAttention needs a multiple of {p_tok} tokens per clip, got {n}")
    inner = mod.heads * 64
    bp = _block_pack(mod, x.device, precision)
    xa = x.reshape(b * n, dim).to(dt).contiguous()
    qk = ops.gemm(xa, bp.w_qkv[: 2 * inner])             # row slices of a contiguous [3*inner, dim] weight are contiguous
    v = ops.gemm(xa, bp.w_qkv[2 * inner:])
    a, _ = ops.attn_temporal(qk, v, b, n // p_tok, p_tok, mod.heads, mod.scale)
    return ops.gemm(a, bp.w_o, bias=bp.b_o, out_dtype=torch.float32).view(b, n, dim)


def _tpack(tr, dev: torch.device, precision: str) -> _TransformerPack:
    cache = tr.__dict__.get("_pack_cache")
    if cache is None:
        cache = _PackCache(pack_transformer)
        object.__setattr__(tr, "_pack_cache", cache)
    return cache.get(tr, dev, precision)


@_inference_only
def transformer_forward(tr, x: torch.Tensor, precision: str = "bf16") -> torch.Tensor:
    """`Transformer.forward`, vivit.py:21-25: x [b, n, dim] -> norm(x_L) [b, n, dim] fp32."""
    dt = _check_precision(precision)
    _check_tokens(x, tr.norm.normalized_shape[0], "Transformer")
    b, n, dim = x.shape
    tp = _tpack(tr, x.device, precision)
    tok = x.reshape(b * n, dim).float().contiguous().clone()
    run_transformer(tp, tok, b, n, dt)
    return ops.layernorm(tok, tp.norm[0], tp.norm[1], torch.float32).view(b, n, dim)


# ------------------------------------------------------------------------------------------------
# ViViT / VanillaTr on block-3 feature maps
# ------------------------------------------------------------------------------------------------
def _features_nhwc(x: torch.Tensor, vit, dt: torch.dtype) -> Tuple[torch.Tensor, int, int]:
    """[b, t, C, h, w] feature maps (the reference's layout, vivit.py:41 / :162) -> patch rows [b*t*h*w, C].  The
    layout change is the reference's own `Rearrange('b t c h w -> b t (h w) c')`; inside `XceptionVidTr` the entry
    flow already produces this layout and nothing is permuted."""
    # ViViT consumes the channels as tokens directly (in_channels == dim); VanillaTr projects them with its patch
    # Linear first (vivit.py:162-167), so its channel count is that layer's in_features, not `dim`
    emb = vit.to_patch_embedding[1] if len(vit.to_patch_embedding) > 1 else None
    chans = emb.in_features if isinstance(emb, torch.nn.Linear) else vit.dim
    if not isinstance(x, torch.Tensor) or x.dim() != 5 or x.shape[2] != chans:
        raise ValueError(f"{type(vit).__name__} expects feature maps [b, t, {chans}, h, w]")
    if not x.is_cuda:
        raise ValueError(f"{type(vit).__name__} (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
    b, t, c, h, w = x.shape
    if t != vit.num_frames or h * w != vit.num_patches:
        raise ValueError(f"feature maps [{t} frames, {h}x{w}] do not match num_frames={vit.num_frames}, "
                         f"num_patches={vit.num_patches}")
    return x.permute(0, 1, 3, 4, 2).to(dt).contiguous().view(b * t * h * w, c), b, t


def vivit_forward_rows(vit, rows: torch.Tensor, b: int, t: int, precision: str, taps: Optional[dict] = None
                       ) -> torch.Tensor:
    """`ViViT.forward`, vivit.py:60-81, from patch rows [b*t*n, dim] (frame-major, row-major patches)."""
    dt = PRECISIONS[precision]
    dev = rows.device
    n, dim = vit.num_patches, vit.dim
    sp = _tpack(vit.space_transformer, dev, precision)
    tp = _tpack(vit.temporal_transformer, dev, precision)
    # space token + patches + pos_embedding per frame (vivit.py:64-66)
    tok = ops.token_build(rows, _f32(vit.space_token.reshape(-1)), _f32(vit.pos_embedding[0]), b * t, n, pos_period=t)
    if taps is not None:
        taps["tokens"] = tok.clone()
    # x[:, 0] of the normalised output: LayerNorm is row-wise, so only the class rows are normalised (vivit.py:25,70-71)
    cls_rows = run_transformer(sp, tok.view(b * t * (n + 1), dim), b * t, n + 1, dt, taps, "space.", cls_only=True)
    cls_rows = ops.layernorm(cls_rows, sp.norm[0], sp.norm[1], torch.float32)
    if taps is not None:
        taps["space_cls"] = cls_rows.clone()
    tok2 = ops.token_build(cls_rows, _f32(vit.temporal_token.reshape(-1)), None, b, t)        # vivit.py:73-74
    head_w, head_b = _f32(vit.mlp_head[1].weight.reshape(-1)), _f32(vit.mlp_head[1].bias)
    if vit.pool == "mean":
        # norm over every frame token, mean over the T+1 tokens of a clip, mlp_head (vivit.py:25, 79-81)
        run_transformer(tp, tok2.view(b * (t + 1), dim), b, t + 1, dt, taps, "temporal.")          # vivit.py:76
        normed = ops.layernorm(tok2.view(b * (t + 1), dim), tp.norm[0], tp.norm[1], torch.float32)
        pooled = ops.layernorm(ops.mean_rows(normed, b, t + 1), *_ln(vit.mlp_head[0]), torch.float32)
        return ops.pool_linear(pooled.view(b, 1, 1, dim), _f32(vit.mlp_head[1].weight), head_b, relu=False)
    cls = run_transformer(tp, tok2.view(b * (t + 1), dim), b, t + 1, dt, taps, "temporal.", cls_only=True)
    # norm, x[:, 0], mlp_head (vivit.py:25, 79-81)
    return ops.head(cls.view(b, 1, 1, dim), tp.norm[0], tp.norm[1], *_ln(vit.mlp_head[0]), head_w, head_b)


def vanilla_forward_rows(vit, rows: torch.Tensor, b: int, t: int, precision: str, taps: Optional[dict] = None
                         ) -> torch.Tensor:
    """`VanillaTr.forward`, vivit.py:179-191, from patch rows [b*t*n, in_channels]."""
    dt = PRECISIONS[precision]
    dev = rows.device
    n, dim = vit.num_patches, vit.dim
    tp = _tpack(vit.transformer, dev, precision)
    emb = vit.to_patch_embedding[1]
    # Linear(patch_dim, dim) on every patch (vivit.py:163), then class token + pos_embedding (vivit.py:183-185)
    e = ops.gemm(rows, emb.weight.detach().to(dt).contiguous(), bias=_f32(emb.bias), out_dtype=torch.float32)
    tok = ops.token_build(e, _f32(vit.cls_token.reshape(-1)), _f32(vit.pos_embedding[0]), b, t * n, pos_period=1)
    del e
    if taps is not None:
        taps["tokens"] = tok.clone()
    seq = t * n + 1
    cls = run_transformer(tp, tok.view(b * seq, dim), b, seq, dt, taps, "", cls_only=True)     # vivit.py:187, 189
    return ops.head(cls.view(b, 1, 1, dim), tp.norm[0], tp.norm[1], *_ln(vit.mlp_head[0]),     # vivit.py:25, 191
                    _f32(vit.mlp_head[1].weight.reshape(-1)), _f32(vit.mlp_head[1].bias))


_ROW_FORWARDS = {"ViViT": vivit_forward_rows, "VanillaTr": vanilla_forward_rows}


@_inference_only
def features_forward(vit, x: torch.Tensor, precision: str = "bf16", taps: Optional[dict] = None) -> torch.Tensor:
    """`ViViT.forward(x)` / `VanillaTr.forward(x)` on feature maps x [b, t, 728, 19, 19] -> logits [b, 1]."""
    dt = _check_precision(precision)
    rows, b, t = _features_nhwc(x, vit, dt)
    return _ROW_FORWARDS[type(vit).__name__](vit, rows, b, t, precision, taps)


def _entry_pack(model, dev: torch.device, precision: str):
    xc = model.xcep.model
    mods = [xc.conv1, xc.bn1, xc.conv2, xc.bn2, xc.block1, xc.block2, xc.block3]
    fp = tuple((t_.data_ptr(), t_._version) for m in mods for t_ in list(m.parameters()) + list(m.buffers()))
    cache = model.__dict__.setdefault("_ablation_entry_packs", {})
    key = (str(dev), precision)
    hit = cache.get(key)
    if hit is None or hit[0] != fp:
        hit = (fp, pack_entry(xc, PRECISIONS[precision]))
        cache[key] = hit
    return hit[1]


@_inference_only
def clip_forward(model, x: torch.Tensor, precision: str = "bf16", taps: Optional[dict] = None) -> torch.Tensor:
    """`XceptionVidTr.forward` (vivit.py:202-208) with `vit` = ViViT or VanillaTr: clips [B, T, 3, H, W] fp32 (or
    decoded uint8 [B, T, H, W, 3]) -> logits [B, 1].  Entry flow as on the ISTVT path; block 3's pooled output is
    already the `b t (h w) c` patch layout, so no permute copy exists."""
    dt = _check_precision(precision)
    vit = model.vit
    is_u8 = isinstance(x, torch.Tensor) and x.dtype == torch.uint8
    if is_u8:
        if x.dim() != 5 or x.shape[4] != 3:
            raise ValueError("uint8 clips must be [B, T, H, W, 3] (decoded frames, channels last)")
        b, t, hh, ww, _ = x.shape
    else:
        if not isinstance(x, torch.Tensor) or x.dim() != 5 or x.shape[2] != 3:
            raise ValueError("ISTVT expects a clip tensor [B, T, 3, H, W]")
        b, t, _, hh, ww = x.shape
    if not x.is_cuda:
        raise ValueError("ISTVT (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
    if t != vit.num_frames:
        raise ValueError(f"clip has {t} frames but the model was built with num_frames={vit.num_frames}")
    ep = _entry_pack(model, x.device, precision)
    if is_u8:
        frames = x.reshape(b * t, hh, ww, 3).contiguous()
        norm = getattr(model, "input_norm", ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)))
    else:
        frames = x.reshape(b * t, 3, hh, ww).float().contiguous()
        norm = None
    body, skip = run_entry_flow(ep, frames, dt, None, norm)
    feat = ops.pool_add(body, skip)                                     # block 3 output, NHWC [b*t, 19, 19, 728]
    n_, fh, fw, c = feat.shape
    if fh * fw != vit.num_patches:
        raise ValueError(f"input {hh}x{ww} reduces to {fh}x{fw} feature maps, the model needs "
                         f"{vit.image_size}x{vit.image_size}")
    if taps is not None:
        taps["block3"] = feat
    return _ROW_FORWARDS[type(vit).__name__](vit, feat.view(n_ * fh * fw, c), b, t, precision, taps)
