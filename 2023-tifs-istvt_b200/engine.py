"""Forward engine of the ISTVT hot path: weight packing + the kernel schedule.

`ISTVTEngine.forward` is what `XceptionVidTr.forward` (reference: network/vivit/vivit.py:202-208) runs.
It owns no arithmetic: every step below is one call into libistvt_b200.so (see include/istvt_b200.h).

Data layout in HBM:
  * feature maps NHWC, activation dtype = bf16 (fp32 in the validation mode);
  * token / residual stream: fp32 [B, F=T+1, P=362, 728]; block-3's pool+add kernel writes straight into it,
    so the reference's `b t c h w -> b t (h w) c` permute, both `cat`s and the `+= pos_embedding`
    (vivit.py:133-142) cost no extra pass;
  * projections are plain row-major [rows, N] and the attention kernels read q/k/v in place.

BatchNorm (eval mode) is folded: scale into the bf16 weights, shift into the GEMM epilogue bias.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

PRECISIONS = {"bf16": torch.bfloat16, "fp32": torch.float32}
_LN2_FOLD = os.environ.get("ISTVT_LN2_FOLD", "1") != "0"
_ROW_PITCH = os.environ.get("ISTVT_ROW_PITCH", "1") != "0"
_SEP_FUSE = os.environ.get("ISTVT_SEP_FUSE", "1") != "0"      # 0: depthwise and pointwise as two kernels everywhere (A/B)


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
PACK_FORMAT = 2      # bump when the layout of a packed tensor changes (v2: K = 728 weights at the aligned row pitch)


def _bn_fold(bn) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


@dataclass
class _SepPack:
    dw: torch.Tensor      # fp32 [3, 3, C]
    pw: torch.Tensor      # act dtype [Cout, Cin], BN scale folded
    bias: torch.Tensor    # fp32 [Cout]


@dataclass
class _BlockPack:
    skip_w: Optional[torch.Tensor]      # None: identity skip (middle flow)
    skip_b: Optional[torch.Tensor]
    seps: List[_SepPack]
    start_with_relu: bool


@dataclass
class _LayerPack:
    ln1: Tuple[torch.Tensor, torch.Tensor]
    w_qk: torch.Tensor
    w_v: torch.Tensor
    w_to: torch.Tensor
    b_to: torch.Tensor
    ln2: Tuple[torch.Tensor, torch.Tensor]
    w_qkv: torch.Tensor
    w_so: torch.Tensor
    b_so: torch.Tensor
    ln3: Tuple[torch.Tensor, torch.Tensor]
    w_1: torch.Tensor
    b_1: torch.Tensor
    w_2: torch.Tensor
    b_2: torch.Tensor
    # LayerNorm 2 folded into to_qkv (bf16 mode): LN(y) W^T = rstd (y (gamma o W)^T - mu rowsum(gamma o W)) + W beta
    w_qkv_f: Optional[torch.Tensor] = None      # bf16 [1536, 728] = gamma o W
    ln2_c: Optional[torch.Tensor] = None        # fp32 [1536]: row sums of the ROUNDED folded weight
    ln2_d: Optional[torch.Tensor] = None        # fp32 [1536]: W beta


def fold_layernorm(weight: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dt: torch.dtype):
    """(gamma o W in `dt`, its fp32 row sums, W beta) for ops.gemm_lnfold; weight [N, K] fp32, gamma / beta [K]."""
    w = weight.detach().float()
    wf = (w * gamma.detach().float()[None, :]).to(dt).contiguous()
    return wf, wf.float().sum(dim=1).contiguous(), (w @ beta.detach().float()).contiguous()


@dataclass
class _EntryPack:
    stem_w: torch.Tensor
    stem_b: torch.Tensor
    conv2_w: torch.Tensor
    conv2_b: torch.Tensor
    blocks: List[_BlockPack]
    u8_key: Optional[tuple] = None        # (mean, std) the uint8-input stem weights below were folded for
    u8_w: Optional[torch.Tensor] = None
    u8_b: Optional[torch.Tensor] = None


@dataclass
class _Pack:
    entry: _EntryPack
    pos_emb: torch.Tensor
    space_token: torch.Tensor
    temporal_token: torch.Tensor
    layers: List[_LayerPack]
    norm: Tuple[torch.Tensor, torch.Tensor]
    head_ln: Tuple[torch.Tensor, torch.Tensor]
    head_w: torch.Tensor
    head_b: torch.Tensor
    fingerprint: tuple = field(default_factory=tuple)


def _f32(t: torch.Tensor) -> torch.Tensor:
    return ops.f32_aligned(t)


def _pack_sep(sep, bn, dt: torch.dtype) -> _SepPack:
    s, b = _bn_fold(bn)
    dw = sep.conv1.weight.detach().float()[:, 0].permute(1, 2, 0).contiguous()            # [C,1,3,3] -> [3,3,C]
    pw = (sep.pointwise.weight.detach().float().flatten(1) * s[:, None]).to(dt).contiguous()
    return _SepPack(dw=dw, pw=pw, bias=b.contiguous())


def _pack_block(block, dt: torch.dtype) -> _BlockPack:
    if block.skip is None:      # middle-flow block: identity skip (xception.py:97-98)
        skip_w = skip_b = None
    else:
        sc, sh = _bn_fold(block.skipbn)
        skip_w = (block.skip.weight.detach().float().flatten(1) * sc[:, None]).to(dt).contiguous()
        skip_b = sh.contiguous()
    seps = []
    mods = list(block.rep)
    for i, m in enumerate(mods):
        if hasattr(m, "pointwise"):
            bn = mods[i + 1]
            s, b = _bn_fold(bn)
            dw = m.conv1.weight.detach().float()[:, 0].permute(1, 2, 0).contiguous()      # [C,1,3,3] -> [3,3,C]
            pw = (m.pointwise.weight.detach().float().flatten(1) * s[:, None]).to(dt).contiguous()
            seps.append(_SepPack(dw=dw, pw=pw, bias=b.contiguous()))
    return _BlockPack(skip_w=skip_w, skip_b=skip_b, seps=seps, start_with_relu=block.start_with_relu)


def pack_entry(xcep, dt: torch.dtype) -> _EntryPack:
    """xcep: the `Xception` parameter tree (network/xception.py)."""
    s1, b1 = _bn_fold(xcep.bn1)
    stem_w = (xcep.conv1.weight.detach().float() * s1[:, None, None, None]).contiguous()   # fp32 [32,3,3,3]
    s2, b2 = _bn_fold(xcep.bn2)
    conv2_w = (xcep.conv2.weight.detach().float() * s2[:, None, None, None]).permute(0, 2, 3, 1).to(dt).contiguous()
    blocks = [_pack_block(b, dt) for b in (xcep.block1, xcep.block2, xcep.block3)]
    return _EntryPack(stem_w=stem_w, stem_b=b1.contiguous(), conv2_w=conv2_w, conv2_b=b2.contiguous(), blocks=blocks)


def fold_input_norm(stem_w: torch.Tensor, stem_b: torch.Tensor, mean, std) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fold the per-channel input normalisation of decoded frames, x_norm = (u8 / 255 - mean_c) / std_c, into the
    stem convolution: conv1 has no padding (xception.py:118), so conv(w, x_norm) + b == conv(w', u8) + b' with
    w'[o, c] = w[o, c] / (255 std_c) and b'[o] = b[o] - sum_c (mean_c / std_c) sum_{ky,kx} w[o, c, ky, kx].
    stem_w: fp32 [32, 3, 3, 3] (BN scale already folded), stem_b: fp32 [32]."""
    mean_t = torch.as_tensor(mean, dtype=torch.float32, device=stem_w.device).reshape(3)
    std_t = torch.as_tensor(std, dtype=torch.float32, device=stem_w.device).reshape(3)
    w = stem_w / (255.0 * std_t)[None, :, None, None]
    b = stem_b - (stem_w.sum(dim=(2, 3)) * (mean_t / std_t)[None, :]).sum(dim=1)
    return w.contiguous(), b.contiguous()


def _on_path_tensors(model) -> List[torch.Tensor]:
    x = model.xcep.model
    mods = [x.conv1, x.bn1, x.conv2, x.bn2, x.block1, x.block2, x.block3, model.vit]
    out: List[torch.Tensor] = []
    for m in mods:
        out += list(m.parameters()) + list(m.buffers())
    return out


def _fingerprint(model) -> tuple:
    return tuple((t.data_ptr(), t._version) for t in _on_path_tensors(model))


def pack_model(model, dt: torch.dtype) -> _Pack:
    vit = model.vit
    ln = lambda m: (_f32(m.weight), _f32(m.bias))
    # bf16 weights sit at the 128-byte aligned row pitch (K = 728 -> 768 elements; ops.pad_rows): a TMA box row of the
    # operand is then one line, not two — 13-25 % on the K = 728 GEMMs (profiles/README.md r6n).  ISTVT_ROW_PITCH=0: A/B.
    pad = ops.pad_rows if (dt == torch.bfloat16 and _ROW_PITCH) else (lambda t: t)
    wt = lambda m: pad(m.weight.detach().to(dt).contiguous())
    layers = []
    for attn_t, attn_s, ff in vit.transformer.layers:
        lp = _LayerPack(
            ln1=ln(attn_t.norm), w_qk=wt(attn_t.fn.to_qk), w_v=wt(attn_t.fn.to_v),
            w_to=wt(attn_t.fn.to_out[0]), b_to=_f32(attn_t.fn.to_out[0].bias),
            ln2=ln(attn_s.norm), w_qkv=wt(attn_s.fn.to_qkv),
            w_so=wt(attn_s.fn.to_out[0]), b_so=_f32(attn_s.fn.to_out[0].bias),
            ln3=ln(ff.norm), w_1=wt(ff.fn.net[0]), b_1=_f32(ff.fn.net[0].bias),
            w_2=wt(ff.fn.net[3]), b_2=_f32(ff.fn.net[3].bias))
        if dt == torch.bfloat16:
            lp.w_qkv_f, lp.ln2_c, lp.ln2_d = fold_layernorm(attn_s.fn.to_qkv.weight, attn_s.norm.weight,
                                                            attn_s.norm.bias, dt)
            lp.w_qkv_f = pad(lp.w_qkv_f)
        layers.append(lp)
    return _Pack(
        entry=pack_entry(model.xcep.model, dt),
        pos_emb=_f32(vit.pos_embedding[0]), space_token=_f32(vit.space_token.reshape(-1)),
        temporal_token=_f32(vit.temporal_token.reshape(-1)),
        layers=layers, norm=ln(vit.transformer.norm), head_ln=ln(vit.mlp_head[0]),
        head_w=_f32(vit.mlp_head[1].weight.reshape(-1)), head_b=_f32(vit.mlp_head[1].bias),
        fingerprint=_fingerprint(model))


# ------------------------------------------------------------------------------------------------
# kernel schedule
# ------------------------------------------------------------------------------------------------
def _run_block(bp: _BlockPack, x: torch.Tensor, taps: Optional[dict], name: str):
    """Xception Block (reference xception.py:91-101): returns (body [n,h,w,C] before the pool, skip)."""
    n, h, w, _ = x.shape
    skip_in = ops.subsample2(x)                                        # gather of the stride-2 1x1 (:57, :94)
    skip = ops.gemm(skip_in, bp.skip_w, bias=bp.skip_b)                # skip conv + skipbn (:94-96)
    y = x
    for i, sp in enumerate(bp.seps):
        relu_in = bp.start_with_relu and i == 0                        # leading ReLU reads the block input (:82-85)
        last = i == len(bp.seps) - 1
        act = ops.ACT_NONE if last else ops.ACT_RELU
        c_in, c_out = y.shape[-1], sp.pw.shape[0]
        if _SEP_FUSE and y.dtype == torch.bfloat16 and ops.sepconv_fused_supported(c_in, c_out, w):
            # depthwise + pointwise + BN (+ ReLU) in one kernel: the depthwise result stays on chip (blocks 1 and 2)
            y = ops.sepconv_fused(y, sp.dw, sp.pw, sp.bias, relu_in, act)
            continue
        d = ops.dwconv3x3(y, sp.dw, relu_in=relu_in)                   # depthwise (:47)
        # pointwise + BN (+ the ReLU that precedes the next separable conv) (:48, :69-75)
        y = ops.gemm(d, sp.pw, bias=sp.bias, act=act)
        y = y.view(n, h, w, -1)
    return y, skip.view(n, skip_in.shape[1], skip_in.shape[2], -1)


def run_entry_flow(ep: _EntryPack, frames: torch.Tensor, dt: torch.dtype, taps: Optional[dict] = None,
                   input_norm=None):
    """frames: fp32 NCHW [n, 3, H, W], or uint8 NHWC [n, H, W, 3] with `input_norm` = (mean, std)
    -> (block-3 body [n, 37, 37, 728], block-3 skip [n, 19, 19, 728])."""
    if frames.dtype == torch.uint8:
        key = (tuple(float(v) for v in input_norm[0]), tuple(float(v) for v in input_norm[1]))
        if ep.u8_key != key:
            ep.u8_w, ep.u8_b = fold_input_norm(ep.stem_w, ep.stem_b, *key)
            ep.u8_key = key
        a = ops.conv_stem_u8(frames, ep.u8_w, ep.u8_b, dt)                             # normalise+conv1+bn1+relu
    else:
        a = ops.conv_stem(frames, ep.stem_w, ep.stem_b, dt)                            # conv1+bn1+relu
    a = ops.conv3x3(a, ep.conv2_w, ep.conv2_b, act=ops.ACT_RELU)                       # conv2+bn2+relu
    if taps is not None:
        taps["stem"] = a
    for bi, bp in enumerate(ep.blocks[:2]):
        body, skip = _run_block(bp, a, taps, f"block{bi + 1}")
        a = ops.pool_add(body, skip)                                                    # maxpool + residual
        if taps is not None:
            taps[f"block{bi + 1}"] = a
    return _run_block(ep.blocks[2], a, taps, "block3")


# ------------------------------------------------------------------------------------------------
# the per-frame Xception baseline: entry flow + middle flow + exit flow + logits
# ------------------------------------------------------------------------------------------------
@dataclass
class _XceptionPack:
    entry: _EntryPack
    middle: List[_BlockPack]       # blocks 4-11
    block12: _BlockPack
    conv3: _SepPack
    conv4: _SepPack
    fc_w: torch.Tensor
    fc_b: torch.Tensor
    fingerprint: tuple = field(default_factory=tuple)


def _xception_fingerprint(xcep) -> tuple:
    return tuple((t.data_ptr(), t._version) for t in list(xcep.parameters()) + list(xcep.buffers()))


def pack_xception(xcep, dt: torch.dtype) -> _XceptionPack:
    fc = xcep.last_linear[-1] if isinstance(xcep.last_linear, torch.nn.Sequential) else xcep.last_linear
    return _XceptionPack(
        entry=pack_entry(xcep, dt),
        middle=[_pack_block(getattr(xcep, f"block{i}"), dt) for i in range(4, 12)],
        block12=_pack_block(xcep.block12, dt),
        conv3=_pack_sep(xcep.conv3, xcep.bn3, dt), conv4=_pack_sep(xcep.conv4, xcep.bn4, dt),
        fc_w=_f32(fc.weight), fc_b=_f32(fc.bias), fingerprint=_xception_fingerprint(xcep))


def _run_identity_block(bp: _BlockPack, x: torch.Tensor) -> torch.Tensor:
    """Middle-flow Block (728 -> 728, 3 separable convs, stride 1, xception.py:130-138): rep(inp) + inp.  The leading
    ReLU is out of place (xception.py:82-85), so the residual adds the un-rectified block input."""
    n, h, w, _ = x.shape
    y = x
    for i, sp in enumerate(bp.seps):
        d = ops.dwconv3x3(y, sp.dw, relu_in=(bp.start_with_relu and i == 0))
        last = i == len(bp.seps) - 1
        y = ops.gemm(d, sp.pw, bias=sp.bias, act=ops.ACT_NONE if last else ops.ACT_RELU).view(n, h, w, -1)
    return ops.add(y, x)


def xception_features(xp: _XceptionPack, frames: torch.Tensor, dt: torch.dtype, input_norm=None,
                      taps: Optional[dict] = None) -> torch.Tensor:
    """Xception.features, xception.py:161-191: frames (fp32 NCHW or uint8 NHWC) -> NHWC [n, 10, 10, 2048] (bn4 output,
    before the ReLU of `logits`)."""
    body, skip = run_entry_flow(xp.entry, frames, dt, taps, input_norm)
    a = ops.pool_add(body, skip)                                                        # block3 output
    if taps is not None:
        taps["block3"] = a
    for i, bp in enumerate(xp.middle):
        a = _run_identity_block(bp, a)                                                  # blocks 4-11
        if taps is not None and i in (0, 7):
            taps[f"block{i + 4}"] = a
    body, skip = _run_block(xp.block12, a, None, "block12")
    a = ops.pool_add(body, skip)                                                        # 19x19 -> 10x10, 1024 ch
    if taps is not None:
        taps["block12"] = a
    n, h, w, _ = a.shape
    d = ops.dwconv3x3(a, xp.conv3.dw, relu_in=False)
    a = ops.gemm(d, xp.conv3.pw, bias=xp.conv3.bias, act=ops.ACT_RELU).view(n, h, w, -1)          # conv3+bn3+relu
    d = ops.dwconv3x3(a, xp.conv4.dw, relu_in=False)
    return ops.gemm(d, xp.conv4.pw, bias=xp.conv4.bias, act=ops.ACT_NONE).view(n, h, w, -1)       # conv4+bn4


class XceptionEngine:
    """Weight-pack cache + schedule of the per-frame Xception forward (`TransferModel('xception')`)."""

    def __init__(self):
        self._packs: Dict[Tuple[str, str], _XceptionPack] = {}

    def pack(self, xcep, dev: torch.device, precision: str) -> _XceptionPack:
        key = (str(dev), precision)
        pk = self._packs.get(key)
        fp = _xception_fingerprint(xcep)
        if pk is None or pk.fingerprint != fp:
            pk = pack_xception(xcep, PRECISIONS[precision])
            self._packs[key] = pk
        return pk

    @torch.no_grad()
    def forward(self, xcep, x: torch.Tensor, precision: str = "bf16", taps: Optional[dict] = None,
                features_only: bool = False) -> torch.Tensor:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if xcep.training:
            raise NotImplementedError("istvt_b200: the per-frame Xception baseline is an inference path (model.eval())")
        is_u8 = isinstance(x, torch.Tensor) and x.dtype == torch.uint8
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or (x.shape[3] if is_u8 else x.shape[1]) != 3:
            raise ValueError("Xception expects frames [N, 3, H, W] (fp32) or decoded frames [N, H, W, 3] (uint8)")
        if not x.is_cuda:
            raise ValueError("Xception (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
        hh, ww = (x.shape[1], x.shape[2]) if is_u8 else (x.shape[2], x.shape[3])
        if min(hh, ww) < 71:
            raise ValueError(f"input {hh}x{ww} is too small for the Xception strides")
        dt = PRECISIONS[precision]
        xp = self.pack(xcep, x.device, precision)
        frames = x.contiguous() if is_u8 else x.float().contiguous()
        norm = getattr(xcep, "input_norm", ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))) if is_u8 else None
        feats = xception_features(xp, frames, dt, norm, taps)
        if features_only:
            return feats
        return ops.pool_linear(feats, xp.fc_w, xp.fc_b, relu=True)                     # logits, xception.py:208-221


def entry_flow_features(xcep, x: torch.Tensor, precision: str = "fp32") -> torch.Tensor:
    """`Xception.low_level_features` in the reference's NCHW fp32 layout (parity / debugging entry)."""
    if xcep.training:
        raise NotImplementedError("istvt_b200: BatchNorm batch-statistics (training-mode) forward is not built yet")
    dt = PRECISIONS[precision]
    ep = pack_entry(xcep, dt)
    body, skip = run_entry_flow(ep, x.contiguous().float(), dt)
    y = ops.pool_add(body, skip)
    return y.float().permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
# packed weights persisted beside the checkpoint (train_CNN.py:998-1011 writes best.pkl)
# ------------------------------------------------------------------------------------------------
def state_digest(model) -> str:
    """sha256 over the bytes of every on-path parameter / buffer, in module order: the key a persisted pack is valid for."""
    import hashlib
    h = hashlib.sha256()
    for t in _on_path_tensors(model):
        c = t.detach().contiguous().cpu()
        h.update(str((tuple(c.shape), str(c.dtype))).encode())
        h.update(c.reshape(-1).view(torch.uint8).numpy().tobytes() if c.numel() else b"")
    return h.hexdigest()


def _pack_to_plain(obj):
    """dataclass tree -> nested dict / list / tuple of CPU tensors (row-pitched views are saved dense, flagged)."""
    from dataclasses import fields, is_dataclass
    if is_dataclass(obj):
        return {"__dc__": type(obj).__name__, **{f.name: _pack_to_plain(getattr(obj, f.name)) for f in fields(obj)
                                                 if f.name != "fingerprint"}}
    if isinstance(obj, torch.Tensor):
        pitched = obj.dim() >= 2 and obj.stride(-2) != obj.shape[-1]
        return {"__t__": obj.detach().contiguous().cpu(), "pitched": pitched}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_pack_to_plain(v) for v in obj)
    return obj


def _pack_from_plain(obj, dev):
    classes = {c.__name__: c for c in (_SepPack, _BlockPack, _LayerPack, _EntryPack, _Pack)}
    if isinstance(obj, dict) and "__dc__" in obj:
        kw = {k: _pack_from_plain(v, dev) for k, v in obj.items() if k != "__dc__"}
        return classes[obj["__dc__"]](**kw)
    if isinstance(obj, dict) and "__t__" in obj:
        t = obj["__t__"].to(dev)
        return ops.pad_rows(t) if obj["pitched"] else t
    if isinstance(obj, (list, tuple)):
        return type(obj)(_pack_from_plain(v, dev) for v in obj)
    return obj


def save_pack(model, path: str, precision: Optional[str] = None) -> str:
    """Write the packed weights of `model` (BatchNorm folded, bf16 casts, LayerNorm-2 fold, aligned row pitch) to
    `path`, keyed by state_digest(model).  Typical use: next to the checkpoint, `save_pack(model, "best.pkl.pack")`."""
    precision = precision or model.precision
    pack = pack_model(model, PRECISIONS[precision])
    blob = {"format": PACK_FORMAT, "precision": precision, "row_pitch": _ROW_PITCH, "digest": state_digest(model),
            "pack": _pack_to_plain(pack)}
    torch.save(blob, path)
    return blob["digest"]


def load_pack(model, path: str) -> bool:
    """Install a persisted pack into the model's engine cache for the model's current device.  Returns False — and
    leaves the engine to pack on the first forward as usual — when the file was written for other weights, another
    format, or another pitch setting; a stale pack is never used."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if blob.get("format") != PACK_FORMAT or blob.get("row_pitch") != _ROW_PITCH or blob.get("digest") != state_digest(model):
        return False
    dev = _on_path_tensors(model)[0].device
    pack = _pack_from_plain(blob["pack"], dev)
    pack.fingerprint = _fingerprint(model)
    eng = model.engine()
    eng._packs[(str(dev), blob["precision"])] = pack
    return True


class ISTVTEngine:
    """Packed-weight cache + forward schedule.  One engine serves a model AND its `nn.DataParallel` replicas
    (train_CNN.py:185-186): packs are cached per (device, precision) and validated against the OWNER's parameters
    (storage pointer + version counter of every on-path tensor), not against the replica's — a replica's parameters
    are fresh broadcast copies on every forward, so a fingerprint of those would force a full re-pack per replica per
    step.  The pack itself is built from the tensors of the module it is asked for (they live on that module's device
    and hold the owner's values)."""

    def __init__(self, model):
        self._owner = weakref.ref(model)
        self._packs: Dict[Tuple[str, str], _Pack] = {}
        self._tensors: Dict[int, List[torch.Tensor]] = {}      # id(module) -> its on-path parameter / buffer objects

    def _fingerprint(self, model) -> tuple:
        """(storage pointer, version counter) of every on-path tensor.  The list of tensor OBJECTS is collected once per
        module (walking the module tree on every forward cost 0.8-1.6 ms of host time — a third of a batch-1 forward):
        in-place updates (optimizer steps, `load_state_dict`, `.data` assignment) change pointer / version and are seen;
        `.to()` / `.cuda()` / `load_state_dict` replace this engine altogether (XceptionVidTr._apply); re-binding an
        attribute to a NEW Parameter object is the one thing that needs `model._engine = None`."""
        owner = self._owner()
        if owner is None or not getattr(model, "_is_replica", False):
            owner = model
        ts = self._tensors.get(id(owner))
        if ts is None:
            ts = self._tensors[id(owner)] = _on_path_tensors(owner)
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _pack(self, model, dev: torch.device, precision: str) -> _Pack:
        key = (str(dev), precision)
        pack = self._packs.get(key)
        fp = self._fingerprint(model)
        if pack is None or pack.fingerprint != fp:
            pack = pack_model(model, PRECISIONS[precision])
            pack.fingerprint = fp
            self._packs[key] = pack
        return pack

    @torch.no_grad()
    def forward(self, model, x: torch.Tensor, precision: str = "bf16", return_attention: bool = False,
                taps: Optional[dict] = None, entry_precision: Optional[str] = None,
                ln2_input_fp32: bool = False):
        """`entry_precision` / `ln2_input_fp32` are error-attribution knobs for tools/precision_probe.py
        (entry flow in another precision than the transformer; temporal-attention output kept fp32 into LN2)."""
        vit = model.vit
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        is_u8 = isinstance(x, torch.Tensor) and x.dtype == torch.uint8
        if is_u8:
            # decoded frames, channels last: the normalisation is folded into the stem (fold_input_norm)
            if x.dim() != 5 or x.shape[4] != 3:
                raise ValueError("uint8 clips must be [B, T, H, W, 3] (decoded frames, channels last)")
        elif not isinstance(x, torch.Tensor) or x.dim() != 5 or x.shape[2] != 3:
            raise ValueError("ISTVT expects a clip tensor [B, T, 3, H, W]")
        if not x.is_cuda:
            raise ValueError("ISTVT (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
        if is_u8:
            b, t, hh, ww, _ = x.shape
        else:
            b, t, _, hh, ww = x.shape
        if t != vit.num_frames:
            raise ValueError(f"clip has {t} frames but the model was built with num_frames={vit.num_frames}")
        if model.training and is_u8:
            raise ValueError("uint8 clips are an inference-path input: normalise to fp32 [B, T, 3, H, W] for training")
        if model.training:
            # train mode without autograd (e.g. under torch.no_grad()): BatchNorm batch statistics, as the reference
            from .train import bridge_trainer
            if return_attention:
                raise ValueError("attention maps are an inference-mode output")
            return bridge_trainer(model, x.device).forward_train(x)[0]
        side = vit.image_size
        feat = lambda s: (((((s - 3) // 2 + 1) - 2) - 1) // 2 + 1)      # stem: 3x3 s2, 3x3 s1; then three s2 stages
        fh, fw = feat(hh), feat(ww)
        for _ in range(2):
            fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
        if (fh, fw) != (side, side):
            raise ValueError(f"input {hh}x{ww} reduces to {fh}x{fw} feature maps, the model needs {side}x{side}")

        dt = PRECISIONS[precision]
        dev = x.device
        pk = self._pack(model, dev, precision)
        heads, dim = vit.heads, vit.dim
        p_tok = vit.num_patches + 1
        f_tok = t + 1
        scale = 64 ** -0.5

        input_norm = None
        if is_u8:
            frames = x.reshape(b * t, hh, ww, 3).contiguous()
            input_norm = getattr(model, "input_norm", ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)))
        else:
            frames = x.reshape(b * t, 3, hh, ww)
            if frames.dtype != torch.float32 or not frames.is_contiguous():
                frames = frames.float().contiguous()

        # ---- Xception entry flow; block 3's pool+add lands in the token buffer ----
        if entry_precision is not None and entry_precision != precision:
            edt = PRECISIONS[entry_precision]
            body, skip = run_entry_flow(self._pack(model, dev, entry_precision).entry, frames, edt, taps, input_norm)
        else:
            body, skip = run_entry_flow(pk.entry, frames, dt, taps, input_norm)
        tokens = torch.empty(b, f_tok, p_tok, dim, dtype=torch.float32, device=dev)
        ops.pool_add_tokens(body, skip, pk.pos_emb, tokens, b, t)
        ops.token_fill(tokens, pk.space_token, pk.temporal_token, pk.pos_emb)
        del body, skip
        if taps is not None:
            taps["tokens"] = tokens.clone()

        attn: List[Tuple[torch.Tensor, torch.Tensor]] = []
        rows = b * f_tok * p_tok
        # GEMM A operands at the aligned row pitch in bf16 mode (see pack_model)
        alloc = ops.empty_rows if (dt == torch.bfloat16 and _ROW_PITCH) else torch.empty
        xn = alloc((b, f_tok, p_tok, dim), dtype=dt, device=dev)
        diff = alloc((b, f_tok, p_tok, dim), dtype=dt, device=dev)
        # After the last block only token (0, 0) of every clip is read (vivit.py:144-148).  Unless intermediates
        # were asked for, the last block therefore runs its temporal attention in full (frame-0 queries need every
        # frame's K / V) and everything after it on the rows that can reach that token: frame-0 rows for the
        # spatial attention, the class-token row for its output projection and the MLP.  bench.py reports the
        # work actually executed.
        prune_last = not return_attention and taps is None and not ln2_input_fp32
        cls_stream = None
        # LayerNorm 2 never runs as a pass in bf16 mode: to_out's epilogue emits the row sums of y1, to_qkv consumes
        # y1 with gamma folded into its weight and applies mu / rstd / beta in its epilogue (ISTVT_LN2_FOLD=0: A/B)
        fold_ln2 = dt == torch.bfloat16 and not ln2_input_fp32 and _LN2_FOLD
        row_stats = torch.empty(rows, (dim + 63) // 64, 2, dtype=torch.float32, device=dev) if fold_ln2 else None
        for li, lp in enumerate(pk.layers):
            # temporal self-subtract attention (module.py:190-208), no residual of its own (vivit.py:99)
            ops.layernorm_diff(tokens, lp.ln1[0], lp.ln1[1], dt, out=(xn, diff))
            qk = ops.gemm(diff, lp.w_qk).view(rows, -1)
            v = ops.gemm(xn, lp.w_v).view(rows, -1)
            at, p_t = ops.attn_temporal(qk, v, b, f_tok, p_tok, heads, scale, want_probs=return_attention)
            if prune_last and li == len(pk.layers) - 1:
                inner = heads * 64
                at0 = ops.gather_rows(at, b, f_tok * p_tok * inner, p_tok, inner, inner)            # frame-0 rows
                yn0 = ops.layernorm(ops.gemm(at0, lp.w_to, bias=lp.b_to, out_dtype=dt), lp.ln2[0], lp.ln2[1], dt)
                as0, _ = ops.attn_spatial(ops.gemm(yn0, lp.w_qkv), b, p_tok, heads, scale)
                as_cls = ops.gather_rows(as0, b, p_tok * inner, 1, inner, inner)                     # query (0, 0)
                x_cls = ops.gather_rows(tokens, b, f_tok * p_tok * dim, 1, dim, dim)                 # fp32 residual rows
                ops.gemm(as_cls, lp.w_so, bias=lp.b_so, residual=x_cls, out=x_cls)
                zn = ops.layernorm(x_cls, lp.ln3[0], lp.ln3[1], dt)
                hid = ops.gemm(zn, lp.w_1, bias=lp.b_1, act=ops.ACT_GELU, out_dtype=dt)
                ops.gemm(hid, lp.w_2, bias=lp.b_2, residual=x_cls, out=x_cls)
                cls_stream = x_cls.view(b, 1, 1, dim)
                break
            # spatial attention (module.py:81-93) + the residual spanning both attentions (vivit.py:99)
            if fold_ln2:
                y1 = ops.gemm_rowstats(at, lp.w_to, lp.b_to, row_stats)
                qkv = ops.gemm_lnfold(y1, lp.w_qkv_f, ops.ln_stats_finalize(row_stats, dim), lp.ln2_c, lp.ln2_d)
            else:
                y1 = ops.gemm(at, lp.w_to, bias=lp.b_to, out_dtype=torch.float32 if ln2_input_fp32 else dt)
                yn = ops.layernorm(y1, lp.ln2[0], lp.ln2[1], dt, out=xn)
                qkv = ops.gemm(yn, lp.w_qkv)
            as_, p_s = ops.attn_spatial(qkv, b * f_tok, p_tok, heads, scale, want_probs=return_attention)
            tok2d = tokens.view(rows, dim)
            ops.gemm(as_, lp.w_so, bias=lp.b_so, residual=tok2d, out=tok2d)
            # MLP (module.py:33-34) + residual (vivit.py:100)
            zn = ops.layernorm(tokens, lp.ln3[0], lp.ln3[1], dt, out=xn)
            hid = ops.gemm(zn, lp.w_1, bias=lp.b_1, act=ops.ACT_GELU, out_dtype=dt).view(rows, -1)
            ops.gemm(hid, lp.w_2, bias=lp.b_2, residual=tok2d, out=tok2d)
            if return_attention:
                attn.append((p_t, p_s.view(b, f_tok, heads, p_tok, p_tok)))
            if taps is not None:
                taps[f"layer{li}"] = tokens.clone()
            del qk, v, at, y1, qkv, as_, hid

        if cls_stream is not None:
            tokens = cls_stream
        logits = ops.head(tokens, pk.norm[0], pk.norm[1], pk.head_ln[0], pk.head_ln[1], pk.head_w, pk.head_b)
        if return_attention:
            return logits, attn
        return logits
