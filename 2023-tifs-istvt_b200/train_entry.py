"""Training-mode forward / backward of the Xception entry flow (network/xception.py:52-101,193-206).

Forward differs from the inference schedule (engine.run_entry_flow) in one way: BatchNorm uses BATCH statistics
(`model.train()`, train_CNN.py:226), so it cannot be folded into the convolution weights — every convolution writes
its raw output, a statistics kernel reduces it per channel, and an apply kernel normalises (+ReLU).  Running
statistics are updated like nn.BatchNorm2d (momentum 0.1, unbiased variance, num_batches_tracked).

Backward per Block (xception.py:91-101), d_out = gradient of `x + skip`:
    main : maxpool'(d_out) -> BN' -> pointwise (dW via split-K GEMM, dX via GEMM with W^T) -> depthwise
           (dW reduction kernel, dX = forward kernel with flipped taps) -> BN'+ReLU' -> pointwise -> depthwise
    skip : BN' -> 1x1 stride-2 conv (dW, dX) ; dX is scattered back to the even pixels
    d_in = main * [ReLU mask of the block input] + scatter(skip)
then conv2 / bn2 / conv1 / bn1 (dense convolutions: weight gradients as GEMMs over im2col^T operands, conv2's data
gradient as the forward implicit-GEMM kernel on the zero-padded gradient with flipped, transposed taps).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict

import torch

from . import ops

BF16 = torch.bfloat16


def _seps(block):
    """[(SeparableConv2d, BatchNorm2d, index of the separable conv in block.rep)] in forward order."""
    mods = list(block.rep)
    return [(m, mods[i + 1], i) for i, m in enumerate(mods) if hasattr(m, "pointwise")]


class EntryFlowTrainer:
    def __init__(self, xcep, mirror=None):
        """mirror: FlatState.bf16_weights() — bf16 views of the parameters by state_dict name, or None."""
        self.x = xcep
        self.mirror = mirror

    def _w(self, weight, name: str):
        from .train import bf16_weight
        return bf16_weight(weight, self.mirror, "xcep.model." + name)

    # ------------------------------------------------------------------ forward
    def _block_fwd(self, blk, xin: torch.Tensor, prefix: str) -> SimpleNamespace:
        n, h, w, cin = xin.shape
        bc = SimpleNamespace(xin=xin)
        bc.skip_in = ops.subsample2(xin)
        ho, wo = bc.skip_in.shape[1:3]
        wsk = self._w(blk.skip.weight, f"{prefix}.skip.weight").flatten(1)
        bc.wskT = ops.transpose(wsk)
        bc.s_raw = ops.gemm(bc.skip_in.view(-1, cin), wsk).view(n, ho, wo, -1)
        bc.s, bc.bn_s = ops.batchnorm_train(bc.s_raw, blk.skipbn, relu=False)
        (sep1, bn_a, i1), (sep2, bn_b, i2) = _seps(blk)
        bc.idx = (i1, i2)
        bc.dw1 = sep1.conv1.weight.detach().float()[:, 0].permute(1, 2, 0).contiguous()      # [3, 3, C]
        bc.dw2 = sep2.conv1.weight.detach().float()[:, 0].permute(1, 2, 0).contiguous()
        pw1 = self._w(sep1.pointwise.weight, f"{prefix}.rep.{i1}.pointwise.weight").flatten(1)
        pw2 = self._w(sep2.pointwise.weight, f"{prefix}.rep.{i2}.pointwise.weight").flatten(1)
        bc.pw1T, bc.pw2T = ops.transpose(pw1), ops.transpose(pw2)
        bc.d1 = ops.dwconv3x3(xin, bc.dw1, relu_in=blk.start_with_relu)
        bc.p1_raw = ops.gemm(bc.d1.view(-1, cin), pw1).view(n, h, w, -1)
        bc.y1, bc.bn_a = ops.batchnorm_train(bc.p1_raw, bn_a, relu=True)
        bc.d2 = ops.dwconv3x3(bc.y1, bc.dw2, relu_in=False)
        bc.p2_raw = ops.gemm(bc.d2.view(-1, bc.d2.shape[-1]), pw2).view(n, h, w, -1)
        bc.y2, bc.bn_b = ops.batchnorm_train(bc.p2_raw, bn_b, relu=False)
        return bc

    def forward(self, frames: torch.Tensor):
        """frames: fp32 NCHW [n, 3, H, W] -> (block-3 body, block-3 skip, context)."""
        x = self.x
        dev = frames.device
        c = SimpleNamespace(frames=frames)
        c.c1 = ops.conv_stem_raw(frames, ops.f32_aligned(x.conv1.weight))
        c.a1, c.bn1 = ops.batchnorm_train(c.c1, x.bn1, relu=True)
        w2 = x.conv2.weight.detach().permute(0, 2, 3, 1).to(BF16).contiguous()                 # [64, 3, 3, 32]
        c.c2 = ops.conv3x3(c.a1, w2, torch.zeros(w2.shape[0], device=dev), act=ops.ACT_NONE)
        a2, c.bn2 = ops.batchnorm_train(c.c2, x.bn2, relu=True)
        xin = a2
        c.blocks = []
        for bi, blk in enumerate((x.block1, x.block2, x.block3)):
            bc = self._block_fwd(blk, xin, f"block{bi + 1}")
            c.blocks.append(bc)
            if bi < 2:
                xin, bc.amax = ops.pool_add_idx(bc.y2, bc.s)
        last = c.blocks[2]
        return last.y2, last.s, c

    # ------------------------------------------------------------------ backward
    def _block_bwd(self, blk, bc, d_out: torch.Tensor, prefix: str, G: Dict[str, torch.Tensor]) -> torch.Tensor:
        n, h, w, c2 = bc.y2.shape
        cin = bc.xin.shape[-1]
        c1 = bc.y1.shape[-1]
        m = n * h * w
        (sep1, bn_a, i1), (sep2, bn_b, i2) = _seps(blk)
        dev = d_out.device

        def dw_grad(xin, dy, name, relu_in):
            tmp = torch.zeros(3, 3, xin.shape[-1], dtype=torch.float32, device=dev)
            ops.dwconv3x3_wgrad(xin, dy, tmp, relu_in)
            G[name].add_(tmp.permute(2, 0, 1).unsqueeze(1))

        # ---- main branch ----
        dy2 = ops.pool_bwd(d_out, bc.amax, h, w)
        dp2 = ops.batchnorm_bwd(dy2, bc.p2_raw, bc.bn_b, G[f"{prefix}.rep.{i2 + 1}.weight"],
                                G[f"{prefix}.rep.{i2 + 1}.bias"], relu=False)
        del dy2
        ops.wgrad(dp2.view(m, c2), bc.d2.view(m, c1), G[f"{prefix}.rep.{i2}.pointwise.weight"].view(c2, c1))
        dd2 = ops.gemm(dp2.view(m, c2), bc.pw2T).view(n, h, w, c1)
        del dp2
        dw_grad(bc.y1, dd2, f"{prefix}.rep.{i2}.conv1.weight", False)
        dy1 = ops.dwconv3x3(dd2, bc.dw2.flip(0, 1).contiguous(), relu_in=False)
        del dd2
        dp1 = ops.batchnorm_bwd(dy1, bc.p1_raw, bc.bn_a, G[f"{prefix}.rep.{i1 + 1}.weight"],
                                G[f"{prefix}.rep.{i1 + 1}.bias"], relu=True)
        del dy1
        ops.wgrad(dp1.view(m, c1), bc.d1.view(m, cin), G[f"{prefix}.rep.{i1}.pointwise.weight"].view(c1, cin))
        dd1 = ops.gemm(dp1.view(m, c1), bc.pw1T).view(n, h, w, cin)
        del dp1
        dw_grad(bc.xin, dd1, f"{prefix}.rep.{i1}.conv1.weight", blk.start_with_relu)
        d_main = ops.dwconv3x3(dd1, bc.dw1.flip(0, 1).contiguous(), relu_in=False)
        del dd1
        # ---- skip branch ----
        ho, wo = bc.s_raw.shape[1:3]
        ms = n * ho * wo
        ds_raw = ops.batchnorm_bwd(d_out, bc.s_raw, bc.bn_s, G[f"{prefix}.skipbn.weight"], G[f"{prefix}.skipbn.bias"],
                                   relu=False)
        ops.wgrad(ds_raw.view(ms, c2), bc.skip_in.view(ms, cin), G[f"{prefix}.skip.weight"].view(c2, cin))
        d_skip_in = ops.gemm(ds_raw.view(ms, c2), bc.wskT).view(n, ho, wo, cin)
        return ops.block_input_grad(d_main, bc.xin, d_skip_in, relu_in=blk.start_with_relu)

    def backward(self, c, amax3: torch.Tensor, g: torch.Tensor, G: Dict[str, torch.Tensor]) -> None:
        """g: fp32 gradient of the token buffer [B, T+1, P, C]; accumulates every entry-flow gradient into G."""
        x = self.x
        dev = g.device
        c.blocks[2].amax = amax3
        d_out = ops.token_grad_gather(g)
        for bi in (2, 1, 0):
            d_out = self._block_bwd((x.block1, x.block2, x.block3)[bi], c.blocks[bi], d_out,
                                    f"xcep.model.block{bi + 1}", G)
            c.blocks[bi] = None
        # ---- bn2 + relu, conv2 ----
        n, h2, w2, co2 = c.c2.shape
        dc2 = ops.batchnorm_bwd(d_out, c.c2, c.bn2, G["xcep.model.bn2.weight"], G["xcep.model.bn2.bias"], relu=True)
        m2 = n * h2 * w2
        ci2 = c.a1.shape[-1]
        tmp = torch.zeros(co2, 9 * ci2, dtype=torch.float32, device=dev)
        ops.gemm_wgrad(ops.transpose(dc2.view(m2, co2)), ops.im2col_t(c.a1), m2, tmp)
        G["xcep.model.conv2.weight"].add_(tmp.view(co2, 3, 3, ci2).permute(0, 3, 1, 2))
        # data gradient: full correlation = valid 3x3 conv of the 2-padded gradient with flipped taps, in/out swapped
        pad = torch.zeros(n, h2 + 4, w2 + 4, co2, dtype=BF16, device=dev)
        pad[:, 2:-2, 2:-2].copy_(dc2)
        del dc2
        w_t = x.conv2.weight.detach().flip(2, 3).permute(1, 2, 3, 0).to(BF16).contiguous()     # [ci, ky, kx, co]
        da1 = ops.conv3x3(pad, w_t, torch.zeros(ci2, device=dev), act=ops.ACT_NONE)
        del pad
        # ---- bn1 + relu, conv1 (no data gradient: the input is the clip) ----
        dc1 = ops.batchnorm_bwd(da1, c.c1, c.bn1, G["xcep.model.bn1.weight"], G["xcep.model.bn1.bias"], relu=True)
        m1 = dc1.numel() // dc1.shape[-1]
        tmp = torch.zeros(32, 32, dtype=torch.float32, device=dev)
        ops.gemm_wgrad(ops.transpose(dc1.view(m1, 32)), ops.im2col_t_stem(c.frames), m1, tmp)
        G["xcep.model.conv1.weight"].add_(tmp[:, :27].reshape(32, 3, 3, 3))
