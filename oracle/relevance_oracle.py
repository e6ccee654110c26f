"""TEST INFRASTRUCTURE — CPU oracle of the relevance pass.  **PARITY UNPINNED.**

The reference's own implementation (`tfe.baselines.ViT.ViT_explanation_generator.LRP`, called at
visualize_rel.py:206,257-262) is absent from its tree and un-vendored: no golden vector, test or
output of the reference exists for this part of the path, so this file cannot be checked against the reference.
It restates, in plain torch with autograd on top of the pinned forward oracle (istvt_oracle.py), the rule the
product implements (2023-tifs-istvt_b200/relevance.py): gradient-weighted attention rollout,
    C_l = mean_heads(relu(dA_l o A_l)),  R = (I + C_L) ... (I + C_1),
spatial map = R_s[frame][0, 1:], temporal map = R_t[position][0, 1:]  (call-site contract visualize_rel.py:257-262).
Only tests/ may import it.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import istvt_oracle as O


def relevance_maps(sd, clips: torch.Tensor, start_layer: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """clips [B, T, 3, H, W] -> (cam_s [B, T, 361], cam_t [B, T, 361], logits [B, 1]); eval mode, fp32.
    `start_layer`: first layer of the rollout (the upstream generator's argument of the same name)."""
    b, t = clips.shape[:2]
    depth = O.num_layers(sd)
    with torch.no_grad():
        feats = O.entry_flow(sd, clips.reshape(b * t, *clips.shape[2:]))
        tokens = O.build_tokens(sd, feats.reshape(b, t, *feats.shape[1:]))
    tokens = tokens.clone().requires_grad_(True)          # something must require grad for the taps to join a graph
    taps = {}
    x = O.transformer(sd, tokens, taps)
    x = x.reshape(b, t + 1, O.TOKENS_PER_FRAME, -1)[:, 0, 0]
    logits = torch.nn.functional.linear(O._ln(sd, "vit.mlp_head.0", x), sd["vit.mlp_head.1.weight"], sd["vit.mlp_head.1.bias"])
    a_s = [taps[f"layer{l}.A_s"] for l in range(depth)]     # [b, h, f, p, p]
    a_t = [taps[f"layer{l}.A_t"] for l in range(depth)]     # [b, h, p, f, f]
    grads = torch.autograd.grad(logits.sum(), a_s + a_t)    # per-clip logits are independent in eval mode
    g_s, g_t = grads[:depth], grads[depth:]
    p, f = O.TOKENS_PER_FRAME, t + 1
    r_s = torch.eye(p).expand(b, f, p, p).clone()
    r_t = torch.eye(f).expand(b, p, f, f).clone()
    for l in range(start_layer, depth):
        c_s = (g_s[l] * a_s[l]).clamp(min=0).mean(dim=1).detach()
        c_t = (g_t[l] * a_t[l]).clamp(min=0).mean(dim=1).detach()
        r_s = r_s + c_s @ r_s
        r_t = r_t + c_t @ r_t
    cam_s = r_s[:, 1:, 0, 1:]                               # [b, T, 361]
    cam_t = r_t[:, 1:, 0, 1:].transpose(1, 2)               # [b, 361, T] -> [b, T, 361]
    return cam_s, cam_t, logits.detach()
