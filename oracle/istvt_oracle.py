"""TEST INFRASTRUCTURE — CPU oracle of the ISTVT forward hot path.

A plain-op fp32 restatement of the reference's algorithm, written against a *state_dict* (reference key
names) so that it needs none of the reference's classes.  Every function cites the
reference lines it follows (paths relative to the reference root).  It exists to check the CUDA path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the
product package (2023-tifs-istvt_b200/) never does.

Pinning: tests/test_oracle.py checks this file (a) against the golden vectors in tests/golden/ that
oracle/make_golden.py produced by running the UNMODIFIED reference modules in the build container, and
(b), where /root/reference is present, against the reference modules directly.
The relevance-propagation part of the path has no reference implementation in the tree:
nothing here covers it — parity for it would be "unpinned".
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5   # nn.BatchNorm2d default (xception.py:119,123,58,69,75)
LN_EPS = 1e-5   # nn.LayerNorm default (module.py:18, vivit.py:89,128)
TOKENS_PER_FRAME = 19 * 19 + 1   # hard-coded in module.py:84,192,197-198 and vivit.py:144


BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


def _bn(sd: SD, prefix: str, x: torch.Tensor, training: bool = False) -> torch.Tensor:
    """BatchNorm2d: eval mode uses the running statistics; training mode (model.train(), train_CNN.py:226) the
    batch statistics, and updates running_mean / running_var (unbiased) / num_batches_tracked in `sd` in place."""
    if training and prefix + ".num_batches_tracked" in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=training, momentum=BN_MOMENTUM, eps=BN_EPS)


def normalise_u8(frames_u8: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """Decoded frames uint8 [B, T, H, W, 3] -> the model input fp32 [B, T, 3, H, W].

    Restates the input transform that precedes the model in the reference: `transforms.ToTensor()` (u8 / 255, HWC ->
    CHW) followed by `transforms.Normalize(mean, std)`.  The transform object itself lives in the absent `dataset`
    package (train_CNN.py:18-21,172-173); xception.py:12-13 documents mean = std = [0.5, 0.5, 0.5] for this
    backbone.  The CUDA path folds the same affine map into the stem convolution."""
    x = frames_u8.to(torch.float32) / 255.0
    m = torch.tensor(mean, dtype=torch.float32).view(1, 1, 1, 1, 3)
    s = torch.tensor(std, dtype=torch.float32).view(1, 1, 1, 1, 3)
    return ((x - m) / s).permute(0, 1, 4, 2, 3).contiguous()


def _sep(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """SeparableConv2d.forward, xception.py:46-49: depthwise 3x3 pad 1 (groups = C) then pointwise 1x1."""
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], None, 1, 1, 1, groups=x.shape[1])
    return F.conv2d(x, sd[prefix + ".pointwise.weight"])


def _block(sd: SD, prefix: str, inp: torch.Tensor, start_with_relu: bool, training: bool = False) -> torch.Tensor:
    """Block.forward, xception.py:91-101, for the entry-flow configuration (reps=2, stride 2, grow_first).

    rep = [ReLU] Sep BN ReLU Sep BN MaxPool(3,2,1) (xception.py:66-89); the leading ReLU is out-of-place
    (xception.py:82-85), so `skip` sees the un-rectified block input.
    """
    i0 = 1 if start_with_relu else 0          # index of the first SeparableConv2d inside rep
    x = F.relu(inp) if start_with_relu else inp
    x = _bn(sd, f"{prefix}.rep.{i0 + 1}", _sep(sd, f"{prefix}.rep.{i0}", x), training)
    x = F.relu(x)
    x = _bn(sd, f"{prefix}.rep.{i0 + 4}", _sep(sd, f"{prefix}.rep.{i0 + 3}", x), training)
    x = F.max_pool2d(x, 3, 2, 1)
    skip = _bn(sd, f"{prefix}.skipbn", F.conv2d(inp, sd[f"{prefix}.skip.weight"], None, 2), training)   # xception.py:94-96
    return x + skip                                                                            # xception.py:100


def entry_flow(sd: SD, frames: torch.Tensor, taps: Optional[dict] = None, prefix: str = "xcep.model",
               training: bool = False) -> torch.Tensor:
    """Xception.low_level_features, xception.py:193-206.  frames [n,3,H,W] -> [n,728,19,19]."""
    x = F.relu(_bn(sd, f"{prefix}.bn1", F.conv2d(frames, sd[f"{prefix}.conv1.weight"], None, 2), training))   # :194-196
    x = F.relu(_bn(sd, f"{prefix}.bn2", F.conv2d(x, sd[f"{prefix}.conv2.weight"]), training))                 # :198-200
    if taps is not None:
        taps["stem"] = x
    x = _block(sd, f"{prefix}.block1", x, start_with_relu=False, training=training)   # xception.py:126
    if taps is not None:
        taps["block1"] = x
    x = _block(sd, f"{prefix}.block2", x, start_with_relu=True, training=training)    # xception.py:127
    if taps is not None:
        taps["block2"] = x
    x = _block(sd, f"{prefix}.block3", x, start_with_relu=True, training=training)    # xception.py:128
    if taps is not None:
        taps["block3"] = x
    return x


# ------------------------------------------------------------------------------------------------
# per-frame Xception baseline (model_selection('xception'), train_CNN.py:924-929): middle + exit flow + logits
# ------------------------------------------------------------------------------------------------
def _block_general(sd: SD, prefix: str, inp: torch.Tensor, reps: int, stride: int) -> torch.Tensor:
    """Block.forward, xception.py:91-101, for the start_with_relu=True blocks 4-12 (eval mode).

    rep = (ReLU Sep BN) x reps [+ MaxPool(3, stride, 1)] (xception.py:66-89; where the channel growth sits —
    grow_first — only changes the weight shapes, which the state_dict carries).  The first ReLU is out of place
    (xception.py:82-85): the skip path sees the un-rectified input.  Identity skip when there is no skip conv
    (xception.py:97-98)."""
    x = inp
    for r in range(reps):
        x = F.relu(x)
        x = _bn(sd, f"{prefix}.rep.{3 * r + 2}", _sep(sd, f"{prefix}.rep.{3 * r + 1}", x))
    if stride != 1:
        x = F.max_pool2d(x, 3, stride, 1)
    if f"{prefix}.skip.weight" in sd:
        skip = _bn(sd, f"{prefix}.skipbn", F.conv2d(inp, sd[f"{prefix}.skip.weight"], None, stride))
    else:
        skip = inp
    return x + skip


def xception_features(sd: SD, frames: torch.Tensor, prefix: str = "model", taps: Optional[dict] = None) -> torch.Tensor:
    """Xception.features, xception.py:161-191.  frames [n,3,H,W] -> [n,2048,h',w'] (bn4 output)."""
    x = entry_flow(sd, frames, taps, prefix=prefix)                       # :162-172
    for i in range(4, 12):                                                 # :173-180, Block(728,728,3,1) :130-138
        x = _block_general(sd, f"{prefix}.block{i}", x, reps=3, stride=1)
        if taps is not None and i in (4, 11):
            taps[f"block{i}"] = x
    x = _block_general(sd, f"{prefix}.block12", x, reps=2, stride=2)       # :181, Block(728,1024,2,2,grow_first=False) :140
    if taps is not None:
        taps["block12"] = x
    x = F.relu(_bn(sd, f"{prefix}.bn3", _sep(sd, f"{prefix}.conv3", x)))    # :183-185
    x = _bn(sd, f"{prefix}.bn4", _sep(sd, f"{prefix}.conv4", x))            # :187-188
    if taps is not None:
        taps["features"] = x
    return x


def xception_forward(sd: SD, frames: torch.Tensor, prefix: str = "model", taps: Optional[dict] = None) -> torch.Tensor:
    """Xception.forward, xception.py:217-220 = features + logits (:208-215: ReLU, adaptive_avg_pool2d(1,1), flatten,
    last_linear = Sequential(Dropout, Linear(2048, classes)) from TransferModel, models_copy.py:40-45; eval mode)."""
    x = F.relu(xception_features(sd, frames, prefix, taps))
    x = F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)
    key = f"{prefix}.last_linear.1" if f"{prefix}.last_linear.1.weight" in sd else f"{prefix}.last_linear"
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def sensitise_xception_(sd: SD, prefix: str = "model", seed: int = 4321, gain: float = 1.85) -> SD:
    """In place, deterministic: default init shrinks the signal by ~1/sqrt(3) per convolution (kaiming_uniform with
    a = sqrt(5)), which would leave the logits of the 36-convolution backbone equal to the classifier bias.  Scale
    every convolution by `gain` and randomise every BatchNorm (gamma, beta, running statistics) so that an error in
    any block reaches the logits."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda shape: torch.rand(shape, generator=g)
    for k in sorted(sd.keys()):
        if not k.startswith(prefix + "."):
            continue
        v = sd[k]
        if k.endswith("running_mean"):
            v.copy_((rnd(v.shape) - 0.5) * 0.2)
        elif k.endswith("running_var"):
            v.copy_(0.5 + rnd(v.shape))
        elif v.dim() == 4:
            v.mul_(gain)
        elif v.dim() == 1 and k.endswith(".weight") and k[: -len(".weight")] + ".running_mean" in sd:
            v.copy_(0.7 + 0.6 * rnd(v.shape))
        elif v.dim() == 1 and k.endswith(".bias") and k[: -len(".bias")] + ".running_mean" in sd:
            v.copy_((rnd(v.shape) - 0.5) * 0.2)
    return sd


def build_tokens(sd: SD, feats: torch.Tensor, prefix: str = "vit") -> torch.Tensor:
    """DSTTr.forward up to the transformer, vivit.py:133-142.  feats [B,T,C,h,w] -> [B,(T+1)*362,C]."""
    b, t, c, h, w = feats.shape
    x = feats.permute(0, 1, 3, 4, 2).reshape(b, t, h * w, c)                     # 'b t c h w -> b t (h w) c' (:115)
    space = sd[f"{prefix}.space_token"].reshape(1, 1, 1, c).expand(b, t, 1, c)   # :136
    x = torch.cat((space, x), dim=2)                                             # :137
    x = x + sd[f"{prefix}.pos_embedding"][:, :, : h * w + 1]                     # :138
    temporal = sd[f"{prefix}.temporal_token"].reshape(1, 1, 1, c).expand(b, 1, h * w + 1, c)   # :139
    x = torch.cat((temporal, x), dim=1)                                          # :140 (no pos-emb on frame 0)
    return x.reshape(b, (t + 1) * (h * w + 1), c)                                # :142


def _ln(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def temporal_attention(sd: SD, prefix: str, xn: torch.Tensor, heads: int = 8, taps: Optional[dict] = None,
                       tag: str = "") -> torch.Tensor:
    """TemporalResidualAttention.forward, module.py:190-208.  xn is the PreNorm output (module.py:21)."""
    b, n, d = xn.shape
    p = TOKENS_PER_FRAME
    f = n // p
    xr = xn.reshape(b, f, p, d)                                                          # :191
    res = torch.cat((xr[:, 0:2], xr[:, 2:] - xr[:, 1:-1]), dim=1).reshape(b, n, d)       # :192-193
    qk = F.linear(res, sd[prefix + ".to_qk.weight"])                                      # :194
    v = F.linear(xn, sd[prefix + ".to_v.weight"])                                         # :195
    q, k = qk.chunk(2, dim=-1)
    split = lambda t: t.reshape(b, f, p, heads, -1).permute(0, 3, 2, 1, 4)               # 'b (t hw) (h d) -> b h hw t d'
    q, k, v = split(q), split(k), split(v)                                                # :196-197
    dots = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)                   # :199, scale :180
    attn = dots.softmax(dim=-1)                                                           # :201
    if taps is not None:
        taps[tag + "A_t"] = attn                                                          # [b, h, hw, f, f]
    out = torch.matmul(attn, v)                                                           # :203
    out = out.permute(0, 3, 2, 1, 4).reshape(b, n, -1)                                    # 'b h hw t d -> b (t hw) (h d)'
    return F.linear(out, sd[prefix + ".to_out.0.weight"], sd[prefix + ".to_out.0.bias"])  # :205


def spatial_attention(sd: SD, prefix: str, xn: torch.Tensor, heads: int = 8, taps: Optional[dict] = None,
                      tag: str = "") -> torch.Tensor:
    """SpatialOnlyAttention.forward, module.py:81-93."""
    b, n, d = xn.shape
    p = TOKENS_PER_FRAME
    f = n // p
    qkv = F.linear(xn, sd[prefix + ".to_qkv.weight"]).chunk(3, dim=-1)                    # :83
    split = lambda t: t.reshape(b, f, p, heads, -1).permute(0, 3, 1, 2, 4)               # 'b (t hw) (h d) -> b h t hw d'
    q, k, v = map(split, qkv)                                                             # :84
    dots = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)                   # :86, scale :72
    attn = dots.softmax(dim=-1)                                                           # :88
    if taps is not None:
        taps[tag + "A_s"] = attn                                                          # [b, h, f, hw, hw]
    out = torch.matmul(attn, v)                                                           # :90
    out = out.permute(0, 2, 3, 1, 4).reshape(b, n, -1)                                    # 'b h t hw d -> b (t hw) (h d)'
    return F.linear(out, sd[prefix + ".to_out.0.weight"], sd[prefix + ".to_out.0.bias"])  # :92


def feed_forward(sd: SD, prefix: str, xn: torch.Tensor) -> torch.Tensor:
    """FeedForward.forward, module.py:27-34: Linear, exact-erf GELU, Linear (dropout p = 0)."""
    h = F.gelu(F.linear(xn, sd[prefix + ".net.0.weight"], sd[prefix + ".net.0.bias"]))
    return F.linear(h, sd[prefix + ".net.3.weight"], sd[prefix + ".net.3.bias"])


def transformer_layer(sd: SD, layer: int, x: torch.Tensor, taps: Optional[dict] = None,
                      prefix: str = "vit.transformer") -> torch.Tensor:
    """One iteration of STTransformer.forward's loop, vivit.py:98-100."""
    lp = f"{prefix}.layers.{layer}"
    tag = f"layer{layer}."
    y = temporal_attention(sd, f"{lp}.0.fn", _ln(sd, f"{lp}.0.norm", x), taps=taps, tag=tag)
    y = spatial_attention(sd, f"{lp}.1.fn", _ln(sd, f"{lp}.1.norm", y), taps=taps, tag=tag)
    x = y + x                                                                             # vivit.py:99
    x = feed_forward(sd, f"{lp}.2.fn", _ln(sd, f"{lp}.2.norm", x)) + x                    # vivit.py:100
    if taps is not None:
        taps[tag + "out"] = x
    return x


def num_layers(sd: SD, prefix: str = "vit.transformer") -> int:
    n = 0
    while f"{prefix}.layers.{n}.0.norm.weight" in sd:
        n += 1
    return n


def transformer(sd: SD, x: torch.Tensor, taps: Optional[dict] = None, prefix: str = "vit.transformer") -> torch.Tensor:
    """STTransformer.forward, vivit.py:97-101."""
    for layer in range(num_layers(sd, prefix)):
        x = transformer_layer(sd, layer, x, taps, prefix)
    return _ln(sd, f"{prefix}.norm", x)                                                   # vivit.py:101


def forward(sd: SD, clips: torch.Tensor, taps: Optional[dict] = None, training: bool = False) -> torch.Tensor:
    """XceptionVidTr.forward, vivit.py:202-208, + DSTTr.forward vivit.py:132-148.  clips [B,T,3,H,W] -> [B,1].
    `training` selects BatchNorm batch statistics (the only train/eval difference on the path: dropout p = 0)."""
    b, t = clips.shape[:2]
    feats = entry_flow(sd, clips.reshape(b * t, *clips.shape[2:]), taps, training=training)   # vivit.py:204-205
    feats = feats.reshape(b, t, *feats.shape[1:])                                         # vivit.py:206
    x = build_tokens(sd, feats)
    if taps is not None:
        taps["tokens"] = x
    x = transformer(sd, x, taps)
    x = x.reshape(b, t + 1, TOKENS_PER_FRAME, -1)[:, 0, 0]                                # vivit.py:144-146
    return F.linear(_ln(sd, "vit.mlp_head.0", x), sd["vit.mlp_head.1.weight"], sd["vit.mlp_head.1.bias"])  # :148


# ------------------------------------------------------------------------------------------------
# training step (train_CNN.py:146-148,196-201,513-533)
# ------------------------------------------------------------------------------------------------
def on_path_parameter_keys(sd: SD) -> List[str]:
    """state_dict keys of the parameters that receive a gradient."""
    keys = []
    for k, v in sd.items():
        if not v.dtype.is_floating_point or k.endswith(("running_mean", "running_var")):
            continue
        if k.startswith("vit.") or any(
                k.startswith(f"xcep.model.{m}") for m in ("conv1.", "bn1.", "conv2.", "bn2.", "block1.", "block2.", "block3.")):
            keys.append(k)
    return keys


def loss_and_grads(sd: SD, clips: torch.Tensor, labels: torch.Tensor, train_entry_flow: bool = True):
    """outputs = model(image); loss = BCEWithLogitsLoss()(outputs.view(-1), labels.float()); loss.backward()
    (train_CNN.py:517,526,532) with the model in train mode.  Returns (loss, logits, {key: grad}); BatchNorm running
    statistics in `sd` are updated in place when the entry flow trains."""
    keys = [k for k in on_path_parameter_keys(sd) if train_entry_flow or k.startswith("vit.")]
    work = dict(sd)
    leaves = {}
    for k in keys:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    logits = forward(work, clips, training=train_entry_flow)
    loss = F.binary_cross_entropy_with_logits(logits.view(-1), labels.float())
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
    if train_entry_flow:
        for k in sd:
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                sd[k] = work[k]
    return loss.detach(), logits.detach(), dict(zip(keys, grads))


def adamw_update(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01) -> None:
    """torch.optim.AdamW (train_CNN.py:199), in place on p, m, v."""
    p.mul_(1 - lr * weight_decay)
    m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    p.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)


# ------------------------------------------------------------------------------------------------
# deterministic "trained-like" weights: at default init every BatchNorm / LayerNorm is
# ~identity and BN/LN folding bugs are invisible; tests randomise them, reproducibly from key names.
# ------------------------------------------------------------------------------------------------
def sensitise_(sd: SD, seed: int = 1234) -> SD:
    """In-place: randomise BN gamma/beta/running stats and LN gamma/beta, shrink pos-emb/tokens to the
    feature scale so that entry-flow errors reach the logit.  Deterministic in (seed, key order)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda shape: torch.rand(shape, generator=g)
    for k in sorted(sd.keys()):
        v = sd[k]
        on_path = k.startswith("vit.") or any(
            k.startswith(f"xcep.model.{m}") for m in ("conv1", "bn1", "conv2", "bn2", "block1.", "block2.", "block3."))
        if not on_path:
            continue
        if k.endswith("running_mean"):
            v.copy_((rnd(v.shape) - 0.5) * 0.2)
        elif k.endswith("running_var"):
            v.copy_(0.5 + rnd(v.shape))
        elif k.endswith("num_batches_tracked"):
            continue
        elif ".norm." in k or "mlp_head.0" in k or "skipbn" in k or ".bn" in k or (
                "xcep.model.block" in k and ".rep." in k and v.dim() == 1):
            if k.endswith("weight"):
                v.copy_(0.75 + 0.5 * rnd(v.shape))
            elif k.endswith("bias"):
                v.copy_((rnd(v.shape) - 0.5) * 0.2)
        elif k in ("vit.pos_embedding", "vit.space_token", "vit.temporal_token"):
            v.mul_(0.05)
    return sd


def fingerprint_indices(numel: int, count: int = 2048, seed: int = 7) -> torch.Tensor:
    """Fixed pseudo-random sample positions used by the golden fixtures for large tensors."""
    g = torch.Generator().manual_seed(seed + numel % 1000003)
    return torch.randint(0, numel, (min(count, numel),), generator=g)
