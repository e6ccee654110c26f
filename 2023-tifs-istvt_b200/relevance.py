"""Relevance maps of the ISTVT transformer (the pass behind visualize_rel.py:206,257-294, BASELINE.json config 4).

The reference obtains them from `tfe.baselines.ViT.ViT_explanation_generator.LRP(model).generate_LRP(image,
method="transformer_attribution", index=0)` — a package that is NOT in the reference tree, so
there is nothing to be parity-pinned against.  What is built here is the published attention-rollout rule of the
same authors' follow-up (Chefer, Gur, Wolf, "Generic Attention-model Explainability", ICCV 2021), which needs only
quantities this framework already produces:

    per layer l and attention kind:   C_l = mean_heads( relu( dA_l o A_l ) )        A = attention probabilities,
                                                                                     dA = d logit / dA
    rollout:                          R = (I + C_L) ... (I + C_1)
    spatial  map of frame f   = row 0 (space class token) of R_s[f],  columns 1..361          -> cam_s [T, 361]
    temporal map of position p = row 0 (temporal class frame) of R_t[p], columns 1..T         -> cam_t [T, 361]

(the LRP variant multiplies dA by an LRP relevance instead of by A; with the reference implementation absent the
choice cannot be verified — DESIGN.md calls this row "parity unpinned").  Data path: one eval-mode CUDA forward that
keeps activations (train.transformer_forward_train), one activation-only backward; the attention backward kernels
accumulate C_l directly (no [B,8,F,362,362] tensor is ever materialised), and because only row 0 of R is read, the
rollout runs as vector-matrix products e_0^T (I + C_L)(I + C_{L-1})... in the order the backward produces C_l.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops
from .engine import pack_entry, run_entry_flow
from .train import BF16, _pack_layers, transformer_backward, transformer_forward_train


class _Rollout:
    def __init__(self, b: int, f: int, p: int, dev, start_layer: int = 0):
        self.b, self.f, self.p = b, f, p
        self.start_layer = start_layer       # rollout over layers start_layer .. L-1 (upstream `start_layer` argument)
        self.cam_s = torch.empty(b * f, p, p, dtype=torch.float32, device=dev)
        self.cam_t = torch.empty(b * p, f, f, dtype=torch.float32, device=dev)
        self.v_s = torch.zeros(b * f, p, dtype=torch.float32, device=dev)
        self.v_t = torch.zeros(b * p, f, dtype=torch.float32, device=dev)
        self.v_s[:, 0] = 1.0
        self.v_t[:, 0] = 1.0

    def spatial_buffer(self) -> torch.Tensor:
        self.cam_s.zero_()
        return self.cam_s

    def temporal_buffer(self) -> torch.Tensor:
        self.cam_t.zero_()
        return self.cam_t

    def layer_done(self, li: int) -> None:
        ops.rollout_row(self.v_s, self.cam_s)
        ops.rollout_row(self.v_t, self.cam_t)


@torch.no_grad()
def relevance_maps(model, x: torch.Tensor, start_layer: int = 0, seed: torch.Tensor = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x: [B, T, 3, H, W] CUDA clips -> (cam_s [B, T, 361], cam_t [B, T, 361], logits [B, 1]); model in eval mode.
    `start_layer`: first transformer layer included in the rollout; `seed`: d(target)/d(logit) per clip (the upstream
    one-hot vector; default 1 = the only logit, class index 0)."""
    if model.training:
        raise ValueError("relevance maps are computed in eval mode (visualize_rel.py:201)")
    if not x.is_cuda:
        raise ValueError("ISTVT (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
    vit = model.vit
    b, t = x.shape[:2]
    if t != vit.num_frames or t + 1 > 48:
        raise ValueError("relevance pass: clip length must equal num_frames (<= 47)")
    f32 = ops.f32_aligned
    frames = x.reshape(b * t, *x.shape[2:]).float().contiguous()
    body, skip = run_entry_flow(pack_entry(model.xcep.model, BF16), frames, BF16)
    pos = f32(vit.pos_embedding[0])
    p = vit.num_patches + 1
    tokens = torch.empty(b, t + 1, p, vit.dim, dtype=torch.float32, device=x.device)
    ops.pool_add_tokens(body, skip, pos, tokens, b, t)
    ops.token_fill(tokens, f32(vit.space_token.reshape(-1)), f32(vit.temporal_token.reshape(-1)), pos)
    del body, skip
    layers = _pack_layers(vit)
    logits, ctxs, x_final, head = transformer_forward_train(vit, layers, tokens)
    depth = len(layers)
    if not 0 <= start_layer < depth:
        raise ValueError(f"start_layer must be in [0, {depth})")
    roll = _Rollout(b, t + 1, p, x.device, start_layer)
    if seed is None:
        dlogits = torch.ones(b, dtype=torch.float32, device=x.device)      # d logit[b, index 0] / d logit[b]
    else:
        dlogits = seed.to(x.device).reshape(b).float().contiguous()
    transformer_backward(vit, layers, ctxs, x_final, head, dlogits, None, relevance=roll)
    cam_s = roll.v_s.view(b, t + 1, p)[:, 1:, 1:].contiguous()                       # [B, T, 361]
    cam_t = roll.v_t.view(b, p, t + 1)[:, 1:, 1:].transpose(1, 2).contiguous()       # [B, 361, T] -> [B, T, 361]
    return cam_s, cam_t, logits


class LRP:
    """Call-site compatible with `attribution_generator = LRP(model)` / `generate_LRP(...)` (visualize_rel.py:206,257):
    for a batch of one clip returns (cam_s, cam_t) as sequences such that `torch.cat(cam_s, 0)` is [T, 361] and
    `torch.cat(cam_t, 0).transpose(0, 1)` is [T, 361], exactly what lines 258-262 index."""

    def __init__(self, model):
        self.model = model

    def generate_LRP(self, x: torch.Tensor, method: str = "transformer_attribution", is_ablation: bool = False,
                     start_layer: int = 0, index=None, **_):
        """The upstream generator's sequence (the public Transformer-Explainability `LRP.generate_LRP`, of which the
        reference's absent `tfe` package is a fork): forward, one-hot vector for `index` (None = the predicted class;
        ISTVT has ONE logit), then `model.relprop(one_hot, method=..., is_ablation=..., start_layer=..., alpha=1)`."""
        if index not in (None, 0):
            raise ValueError("ISTVT has one logit: index must be 0")
        if x.shape[0] != 1:
            raise ValueError("generate_LRP follows the reference call site (batch 1); use relevance_maps() for batches")
        one_hot = torch.ones(1, 1, dtype=torch.float32, device=x.device)
        return self.model.relprop(one_hot, method=method, is_ablation=is_ablation, start_layer=start_layer, alpha=1,
                                  clips=x)


RELPROP_METHODS = ("transformer_attribution", "attn_grad_rollout")


def relprop(model, cam: torch.Tensor = None, method: str = "transformer_attribution", is_ablation: bool = False,
            start_layer: int = 0, alpha: float = 1, clips: torch.Tensor = None):
    """`model.relprop(one_hot, method=..., is_ablation=..., start_layer=..., alpha=1)` of the upstream convention
    (SURVEY.md section 8(b)): returns (cam_s, cam_t) as sequences such that `torch.cat(cam_s, 0)` is [T, 361] and
    `torch.cat(cam_t, 0).transpose(0, 1)` is [T, 361] — exactly what visualize_rel.py:258-262 indexes.  `clips`: the
    batch-1 input; without it the clip of the model's last eval-mode forward is used (`model.keep_relprop_input`)."""
    if method not in RELPROP_METHODS:
        raise NotImplementedError(f"relevance method {method!r} is not built (available: {RELPROP_METHODS})")
    if is_ablation:
        raise NotImplementedError("is_ablation is an option of the absent upstream package; not built")
    if alpha != 1:
        raise NotImplementedError("only alpha = 1 (the reference call sites' value) is built")
    if clips is None:
        clips = getattr(model, "_relprop_clips", None)
        if clips is None:
            raise ValueError("relprop needs the clip: pass clips=..., or set model.keep_relprop_input = True before the "
                             "forward whose decision is to be explained")
    if clips.shape[0] != 1:
        raise ValueError("relprop follows the reference call site (batch 1); use relevance_maps() for batches")
    cam_s, cam_t, _ = relevance_maps(model, clips, start_layer=start_layer, seed=cam)
    seq_s: List[torch.Tensor] = [cam_s[0, i:i + 1] for i in range(cam_s.shape[1])]              # T x [1, 361]
    seq_t: List[torch.Tensor] = [cam_t[0, :, j:j + 1].t() for j in range(cam_t.shape[2])]       # 361 x [1, T]
    return seq_s, seq_t
