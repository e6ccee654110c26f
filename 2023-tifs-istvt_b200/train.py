"""Training step of the ISTVT hot path: forward that keeps activations, hand-written backward, AdamW, DP all-reduce.

Mirrors the ISTVT branch of the reference's training loop (train_CNN.py:146-148,196-201,513-533):

    optimizer.zero_grad(); outputs = model(image)
    loss = nn.BCEWithLogitsLoss()(outputs.view(-1), labels.float()); loss.backward(); optimizer.step()

with `optim.AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay)` and, on several GPUs, the gradient reduction
that `nn.DataParallel` performs (train_CNN.py:185-186) done as ONE NCCL all-reduce of a flat fp32 gradient buffer
(one process per GPU, BatchNorm statistics rank-local exactly like DataParallel replicas).

Every arithmetic step is a kernel of libistvt_b200.so (see ops.py); torch supplies memory, streams, the
loss on the [B] logits (the caller-side line train_CNN.py:526) and `torch.distributed`.  bf16 mode only.

Gradient flow per spatial-temporal block (forward: engine.py / vivit.py:97-100), g = dL/d(residual stream), fp32:
    MLP      : dW2,db2 ; dhid = g W2 ; dhpre = dhid * gelu'(hpre) ; dW1,db1 ; dzn = dhpre W1 ; g += LN3'(dzn)
    spatial  : dWso,dbso ; das = g Wso ; dqkv = attn_s'(das) ; dWqkv ; dyn = dqkv Wqkv ; dy1 = LN2'(dyn)
    temporal : dWto,dbto ; dat = dy1 Wto ; (dqk, dv) = attn_t'(dat) ; dWqk, dWv ;
               g += LN1'( dv Wv  +  self-subtract'(dqk Wqk) )
Weight gradients are split-K tcgen05 GEMMs that read dY and X in place as MN-major operand tiles.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .engine import pack_entry, run_entry_flow

BF16 = torch.bfloat16
# ISTVT_MLP_FUSE=1: GELU (+ pre-activation) in ff1's epilogue and gelu' in the ff2 data-gradient epilogue instead of the
# stand-alone GELU passes.  Measured on the C3 step (profiles/README.md r6q): the passes cost 4.4 + 5.9 ms, the fused
# epilogues add 8.8 ms to the GEMMs (twice the TMA store boxes; per-lane loads of the pre-activation) — 208.5 -> 207.0 ms,
# not enough to make it the default.
_MLP_FUSE = os.environ.get("ISTVT_MLP_FUSE", "0") != "0"


# ------------------------------------------------------------------------------------------------
# parameters on the path, flat buffers
# ------------------------------------------------------------------------------------------------
def _named_parameters(module):
    """`module.named_parameters()` that also works on an `nn.DataParallel` replica: `replicate()` empties
    `_parameters` and keeps the broadcast copies (non-leaf tensors whose autograd history reduces the gradients back
    to the source device) as plain attributes, listed in `_former_parameters` — same order as named_parameters()."""
    if not getattr(module, "_is_replica", False):
        yield from module.named_parameters()
        return
    for mname, m in module.named_modules():
        for k, p in getattr(m, "_former_parameters", {}).items():
            if p is not None:
                yield (f"{mname}.{k}" if mname else k), p


def on_path_named_parameters(model, train_entry_flow: bool) -> List[Tuple[str, torch.Tensor]]:
    """Parameters that receive a gradient (117 Xception-tail tensors never do)."""
    out = []
    if train_entry_flow:
        x = model.xcep.model
        for name in ("conv1", "bn1", "conv2", "bn2", "block1", "block2", "block3"):
            for n, p in _named_parameters(getattr(x, name)):
                out.append((f"xcep.model.{name}.{n}", p))
    for n, p in _named_parameters(model.vit):
        out.append((f"vit.{n}", p))
    return out


class FlatState:
    """fp32 parameters / gradients / Adam moments of the path in four flat buffers; every `nn.Parameter.data`
    becomes a view into `params`, so state_dict(), the packing code and external optimizers keep working.

    `grads_only` (an `nn.DataParallel` replica: its parameters are per-step broadcast copies owned by autograd, the
    optimizer runs on the source module): only the flat gradient buffer and its per-parameter views exist."""

    def __init__(self, model, train_entry_flow: bool, grads_only: bool = False):
        named = on_path_named_parameters(model, train_entry_flow)
        dev = named[0][1].device
        sizes = [(p.numel() + 7) // 8 * 8 for _, p in named]       # slots 16-byte aligned in fp32 AND in the bf16 mirror
        total = sum(sizes)
        self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params = self.exp_avg = self.exp_avg_sq = None
        if not grads_only:
            self.params = torch.zeros_like(self.grads)
            self.exp_avg = torch.zeros_like(self.grads)
            self.exp_avg_sq = torch.zeros_like(self.grads)
        self.grad: Dict[str, torch.Tensor] = {}
        self.names = [n for n, _ in named]
        off = 0
        for (name, p), sz in zip(named, sizes):
            if not grads_only:
                view = self.params[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
            self.grad[name] = self.grads[off:off + p.numel()].view(p.shape)
            off += sz
        self.step = 0
        self._offsets = {}
        off = 0
        for (name, p), sz in zip(named, sizes):
            self._offsets[name] = (off, p.shape)
            off += sz
        self.params_bf16: Optional[torch.Tensor] = None
        # gradient buckets of the data-parallel reduction: the entry flow's slots come first, the transformer's last
        self.vit_offset = next((o for n, (o, _) in self._offsets.items() if n.startswith("vit.")), 0)

    def bf16_weights(self) -> Optional[Dict[str, torch.Tensor]]:
        """bf16 mirror of every on-path parameter, refreshed by ONE cast kernel over the flat master buffer (the
        optimizer moves the weights every step); views share the fp32 slots' offsets.  None for a replica."""
        if self.params is None:
            return None
        if self.params_bf16 is None:
            self.params_bf16 = torch.empty(self.params.shape, dtype=BF16, device=self.params.device)
        ops.cast_bf16(self.params, out=self.params_bf16)
        return {n: self.params_bf16[o:o + torch.Size(shape).numel()].view(shape) for n, (o, shape) in self._offsets.items()}

    def attach_grads(self, model, train_entry_flow: bool) -> None:
        """Expose the flat gradient views as `.grad` (for inspection / external optimizers)."""
        for name, p in on_path_named_parameters(model, train_entry_flow):
            p.grad = self.grad[name]


# ------------------------------------------------------------------------------------------------
# weight packs for training (bf16 forward weights + their transposes for the data-gradient GEMMs)
# ------------------------------------------------------------------------------------------------
def bf16_weight(w: torch.Tensor, mirror: Optional[Dict[str, torch.Tensor]], name: str) -> torch.Tensor:
    """bf16 copy of an fp32 weight: the view of the flat mirror when there is one, else one cast kernel."""
    if mirror is not None and name in mirror:
        return mirror[name]
    return ops.cast_bf16(ops.f32_aligned(w))


def _pack_layers(vit, mirror: Optional[Dict[str, torch.Tensor]] = None) -> List[SimpleNamespace]:
    """Per-layer bf16 weights [N, K] and their transposes [K, N] (B operand of the data-gradient GEMMs); `mirror` =
    FlatState.bf16_weights().  The transposes are produced by istvt_transpose_colsum — no ATen kernel on the path."""
    layers = []
    f32 = ops.f32_aligned
    for li, (attn_t, attn_s, ff) in enumerate(vit.transformer.layers):
        L = SimpleNamespace()
        pre = f"vit.transformer.layers.{li}"
        for key, lin, name in (("qk", attn_t.fn.to_qk, f"{pre}.0.fn.to_qk"), ("v", attn_t.fn.to_v, f"{pre}.0.fn.to_v"),
                               ("to", attn_t.fn.to_out[0], f"{pre}.0.fn.to_out.0"), ("qkv", attn_s.fn.to_qkv, f"{pre}.1.fn.to_qkv"),
                               ("so", attn_s.fn.to_out[0], f"{pre}.1.fn.to_out.0"), ("1", ff.fn.net[0], f"{pre}.2.fn.net.0"),
                               ("2", ff.fn.net[3], f"{pre}.2.fn.net.3")):
            w = bf16_weight(lin.weight, mirror, name + ".weight")
            # GEMM operands whose K rows are not 64-byte aligned (K = 728) sit at the 128-byte aligned pitch: the forward
            # weight as a pitched cast of the fp32 master, the transposed weight through the transpose kernel's ldo
            wf = w
            if ops.row_pitch(w.shape[1]) != w.shape[1]:
                wf = ops.cast_bf16(f32(lin.weight), out=ops.empty_rows(tuple(w.shape), BF16, w.device))
            setattr(L, "w_" + key, wf)
            setattr(L, "wT_" + key, ops.transpose(w, aligned=True))
            if lin.bias is not None:
                setattr(L, "b_" + key, f32(lin.bias))
        L.ln1 = (f32(attn_t.norm.weight), f32(attn_t.norm.bias))
        L.ln2 = (f32(attn_s.norm.weight), f32(attn_s.norm.bias))
        L.ln3 = (f32(ff.norm.weight), f32(ff.norm.bias))
        layers.append(L)
    return layers


def _layer_grad_names(i: int) -> Dict[str, str]:
    p = f"vit.transformer.layers.{i}"
    return {
        "ln1_w": f"{p}.0.norm.weight", "ln1_b": f"{p}.0.norm.bias",
        "w_qk": f"{p}.0.fn.to_qk.weight", "w_v": f"{p}.0.fn.to_v.weight",
        "w_to": f"{p}.0.fn.to_out.0.weight", "b_to": f"{p}.0.fn.to_out.0.bias",
        "ln2_w": f"{p}.1.norm.weight", "ln2_b": f"{p}.1.norm.bias",
        "w_qkv": f"{p}.1.fn.to_qkv.weight",
        "w_so": f"{p}.1.fn.to_out.0.weight", "b_so": f"{p}.1.fn.to_out.0.bias",
        "ln3_w": f"{p}.2.norm.weight", "ln3_b": f"{p}.2.norm.bias",
        "w_1": f"{p}.2.fn.net.0.weight", "b_1": f"{p}.2.fn.net.0.bias",
        "w_2": f"{p}.2.fn.net.3.weight", "b_2": f"{p}.2.fn.net.3.bias",
    }


# ------------------------------------------------------------------------------------------------
# transformer: forward keeping activations, backward
# ------------------------------------------------------------------------------------------------
def transformer_forward_train(vit, layers: List[SimpleNamespace], tokens: torch.Tensor):
    """tokens: fp32 [B, F, P, D] (not modified).  Returns (logits [B, 1], per-layer contexts, final stream)."""
    b, f, p, d = tokens.shape
    rows = b * f * p
    heads = vit.heads
    scale = 64 ** -0.5
    f32 = ops.f32_aligned
    act = lambda: ops.empty_rows((rows, d), BF16, tokens.device)
    act4 = lambda: ops.empty_rows((b, f, p, d), BF16, tokens.device)
    ctxs = []
    x = tokens
    for L in layers:
        c = SimpleNamespace()
        c.x0 = x
        # every bf16 [rows, 728] activation is a GEMM / weight-gradient operand: 128-byte aligned row pitch (act())
        c.xn, c.diff = ops.layernorm_diff(x, L.ln1[0], L.ln1[1], BF16, out=(act4(), act4()))
        c.xn, c.diff = c.xn.view(rows, d), c.diff.view(rows, d)
        c.qk = ops.gemm(c.diff, L.w_qk)
        c.v = ops.gemm(c.xn, L.w_v)
        c.at, _ = ops.attn_temporal(c.qk, c.v, b, f, p, heads, scale)
        c.y1 = ops.gemm(c.at, L.w_to, bias=L.b_to, out=act())
        c.yn = ops.layernorm(c.y1, L.ln2[0], L.ln2[1], BF16, out=act())
        c.qkv = ops.gemm(c.yn, L.w_qkv)
        c.as_, c.lse = ops.attn_spatial_lse(c.qkv, b * f, p, heads, scale)
        x1 = torch.empty_like(x)
        ops.gemm(c.as_, L.w_so, bias=L.b_so, residual=x.view(rows, d), out=x1.view(rows, d))
        c.x1 = x1
        c.zn = ops.layernorm(x1.view(rows, d), L.ln3[0], L.ln3[1], BF16, out=act())
        if _MLP_FUSE:      # GELU in ff1's epilogue, both the activated value and the pre-activation are written
            c.hid, c.hpre = ops.gemm_act_dual(c.zn, L.w_1, L.b_1, ops.ACT_GELU)
        else:
            c.hpre = ops.gemm(c.zn, L.w_1, bias=L.b_1, out_dtype=BF16)
            c.hid = ops.gelu(c.hpre)
        x2 = torch.empty_like(x)
        ops.gemm(c.hid, L.w_2, bias=L.b_2, residual=x1.view(rows, d), out=x2.view(rows, d))
        x = x2
        ctxs.append(c)
    head = SimpleNamespace(norm=(f32(vit.transformer.norm.weight), f32(vit.transformer.norm.bias)),
                           ln=(f32(vit.mlp_head[0].weight), f32(vit.mlp_head[0].bias)),
                           w=f32(vit.mlp_head[1].weight.reshape(-1)), b=f32(vit.mlp_head[1].bias))
    logits = ops.head(x, head.norm[0], head.norm[1], head.ln[0], head.ln[1], head.w, head.b)
    return logits, ctxs, x, head


class _ScratchGrads(dict):
    """Stand-in for the gradient dictionary when only activation gradients are wanted (relevance pass): the
    LayerNorm / head kernels still need somewhere to put dgamma / dbeta."""

    def __init__(self, dev):
        super().__init__()
        self.dev = dev

    def __missing__(self, key):
        t = torch.zeros(4096, dtype=torch.float32, device=self.dev)
        self[key] = t
        return t


def transformer_backward(vit, layers, ctxs, x_final: torch.Tensor, head, dlogits: torch.Tensor,
                         G: Optional[Dict[str, torch.Tensor]], relevance=None) -> torch.Tensor:
    """Accumulates every vit.transformer / vit.mlp_head gradient into G and returns g = dL/d(tokens), fp32.
    G = None: activation gradients only (no weight-gradient GEMMs).  `relevance` (relevance.py): per-layer hook that
    receives the head-averaged relu(dA o A) of both attentions."""
    b, f, p, d = x_final.shape
    rows = b * f * p
    heads = vit.heads
    scale = 64 ** -0.5
    wgrad = G is not None
    if G is None:
        G = _ScratchGrads(x_final.device)
    g = torch.zeros_like(x_final)
    g2 = g.view(rows, d)
    ops.head_bwd(x_final, dlogits.reshape(-1).float().contiguous(), head.norm[0], head.norm[1], head.ln[0], head.ln[1],
                 head.w, g, G["vit.transformer.norm.weight"], G["vit.transformer.norm.bias"],
                 G["vit.mlp_head.0.weight"], G["vit.mlp_head.0.bias"], G["vit.mlp_head.1.weight"].view(-1)[:d],
                 G["vit.mlp_head.1.bias"])
    act = lambda: ops.empty_rows((rows, d), BF16, g.device)      # bf16 [rows, 728] operands at the aligned row pitch
    g_bf = ops.cast_bf16(g2, out=act())
    scratch = torch.empty(rows, heads * 64, dtype=torch.float32, device=g.device)
    # Bias gradients = column sums of the output gradients.  They are taken by the kernel that PRODUCES each gradient
    # (layernorm_bwd for g / dy1, gelu_bwd for dhpre) instead of by a column-sum pass over what it has just written;
    # only the very first g (from the head) needs the stand-alone kernel.
    if wgrad:
        ops.colsum(g_bf, G[_layer_grad_names(len(layers) - 1)["b_2"]])
    for li in range(len(layers) - 1, -1, -1):
        L, c = layers[li], ctxs[li]
        N = {k: G[v] for k, v in _layer_grad_names(li).items()}
        if not wgrad:
            if relevance is not None and li < getattr(relevance, "start_layer", 0):
                break                        # the rollout starts at `start_layer`: nothing below it is read
            _transformer_layer_backward_acts(L, c, N, g2, g_bf, scratch, b, f, p, d, heads, scale, relevance, li)
            ctxs[li] = None
            continue
        # ---- MLP (module.py:27-34) ----
        ops.wgrad(g_bf, c.hid, N["w_2"])                    # db_2: column sums of g, taken where g was produced
        if _MLP_FUSE:      # gelu'(hpre) applied in the data-gradient GEMM's epilogue
            dhpre = ops.gemm_dgelu(g_bf, L.wT_2, c.hpre)
            ops.colsum(dhpre, N["b_1"])
        else:
            dhpre = ops.gelu_bwd(ops.gemm(g_bf, L.wT_2), c.hpre, colsum=N["b_1"])
        ops.wgrad(dhpre, c.zn, N["w_1"])
        dzn = ops.gemm(dhpre, L.wT_1, out=act())
        del dhpre
        ops.layernorm_bwd(dzn, c.x1.view(rows, d), L.ln3[0], N["ln3_w"], N["ln3_b"], g_accum=g2, g_bf16=g_bf,
                          out_colsum=N["b_so"])
        del dzn
        # ---- spatial attention (module.py:81-93) ----
        ops.wgrad(g_bf, c.as_, N["w_so"])
        das = ops.gemm(g_bf, L.wT_so)
        dqkv = ops.attn_spatial_bwd(c.qkv, c.as_, das, c.lse, b * f, p, heads, scale, scratch)
        del das
        ops.wgrad(dqkv, c.yn, N["w_qkv"])
        dyn = ops.gemm(dqkv, L.wT_qkv, out=act())
        del dqkv
        dy1 = ops.layernorm_bwd(dyn, c.y1, L.ln2[0], N["ln2_w"], N["ln2_b"], out_colsum=N["b_to"])
        del dyn
        # ---- temporal self-subtract attention (module.py:190-208) ----
        ops.wgrad(dy1, c.at, N["w_to"])
        dat = ops.gemm(dy1, L.wT_to)
        del dy1
        dqk, dv = ops.attn_temporal_bwd(c.qk, c.v, dat, b, f, p, heads, scale)
        del dat
        ops.wgrad(dqk, c.diff, N["w_qk"])
        ops.wgrad(dv, c.xn, N["w_v"])
        ddiff = ops.gemm(dqk, L.wT_qk, out=act())
        dxn_v = ops.gemm(dv, L.wT_v, out=act())
        del dqk, dv
        ops.layernorm_bwd(dxn_v, c.x0.view(rows, d), L.ln1[0], N["ln1_w"], N["ln1_b"], g_accum=g2, g_bf16=g_bf,
                          dy2=ddiff, frames=f, tokens_per_frame=p,
                          out_colsum=G[_layer_grad_names(li - 1)["b_2"]] if li > 0 else None)   # db_2 of the layer below
        del ddiff, dxn_v
        ctxs[li] = None     # release this layer's activations
    return g


def _transformer_layer_backward_acts(L, c, N, g2, g_bf, scratch, b, f, p, d, heads, scale, relevance, li) -> None:
    """One block's backward without the weight-gradient GEMMs (same data path as transformer_backward)."""
    rows = b * f * p
    act = lambda: ops.empty_rows((rows, d), BF16, g2.device)
    dhpre = ops.gemm_dgelu(g_bf, L.wT_2, c.hpre) if _MLP_FUSE else ops.gelu_bwd(ops.gemm(g_bf, L.wT_2), c.hpre)
    dzn = ops.gemm(dhpre, L.wT_1, out=act())
    del dhpre
    ops.layernorm_bwd(dzn, c.x1.view(rows, d), L.ln3[0], N["ln3_w"], N["ln3_b"], g_accum=g2, g_bf16=g_bf)
    del dzn
    das = ops.gemm(g_bf, L.wT_so)
    cam_s = relevance.spatial_buffer() if relevance is not None else None
    dqkv = ops.attn_spatial_bwd(c.qkv, c.as_, das, c.lse, b * f, p, heads, scale, scratch, cam=cam_s)
    del das
    dy1 = ops.layernorm_bwd(ops.gemm(dqkv, L.wT_qkv, out=act()), c.y1, L.ln2[0], N["ln2_w"], N["ln2_b"])
    del dqkv
    dat = ops.gemm(dy1, L.wT_to)
    del dy1
    cam_t = relevance.temporal_buffer() if relevance is not None else None
    dqk, dv = ops.attn_temporal_bwd(c.qk, c.v, dat, b, f, p, heads, scale, cam=cam_t)
    del dat
    ddiff = ops.gemm(dqk, L.wT_qk, out=act())
    dxn_v = ops.gemm(dv, L.wT_v, out=act())
    del dqk, dv
    ops.layernorm_bwd(dxn_v, c.x0.view(rows, d), L.ln1[0], N["ln1_w"], N["ln1_b"], g_accum=g2, g_bf16=g_bf,
                      dy2=ddiff, frames=f, tokens_per_frame=p)
    if relevance is not None:
        relevance.layer_done(li)


# ------------------------------------------------------------------------------------------------
# the trainer
# ------------------------------------------------------------------------------------------------
class Trainer:
    """fwd + BCE-with-logits + bwd + (all-reduce) + AdamW for `XceptionVidTr` on one GPU per process.

    `train_entry_flow=False` keeps the Xception entry flow in eval mode (running BatchNorm statistics, no
    gradients): fine-tuning of the spatial-temporal transformer only.
    """

    def __init__(self, model, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 train_entry_flow: bool = True, process_group=None):
        if model.precision != "bf16":
            raise ValueError("istvt_b200 training runs in bf16 mode (fp32 master weights, fp32 gradients)")
        # an nn.DataParallel replica (train_CNN.py:185-186): forward / backward only, gradients flow back to the
        # source module through autograd's Broadcast node; the optimizer belongs to the source module
        self.replica = bool(getattr(model, "_is_replica", False))
        if on_path_named_parameters(model, True)[0][1].device.type != "cuda":
            raise ValueError("move the model to a CUDA device first: there is no CPU training path")
        if model.vit.num_frames + 1 > 48:
            raise NotImplementedError("training covers clips of up to 47 frames (temporal-attention backward kernels)")
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.train_entry_flow = train_entry_flow
        self.pg = process_group
        self.state = FlatState(model, train_entry_flow, grads_only=self.replica)
        if not self.replica:
            model._engine = None

    # -- pieces (also used by the autograd.Function wrapper) --
    def forward_train(self, x: torch.Tensor):
        model, vit = self.model, self.model.vit
        b, t = x.shape[:2]
        if t != vit.num_frames:
            raise ValueError(f"clip has {t} frames but the model was built with num_frames={vit.num_frames}")
        frames = x.reshape(b * t, *x.shape[2:]).float().contiguous()
        f32 = ops.f32_aligned
        pos = f32(vit.pos_embedding[0])
        tokens = torch.empty(b, t + 1, vit.num_patches + 1, vit.dim, dtype=torch.float32, device=x.device)
        mirror = self.state.bf16_weights()
        entry = None
        if self.train_entry_flow:
            from .train_entry import EntryFlowTrainer
            entry = EntryFlowTrainer(model.xcep.model, mirror)   # stateless: binds the module tree this forward ran on
            body, skip, ectx = entry.forward(frames)
            _, amax3 = ops.pool_add_idx(body, skip, tokens=tokens, pos_emb=pos, t_frames=t)
        else:
            body, skip = run_entry_flow(pack_entry(model.xcep.model, BF16), frames, BF16)
            ectx, amax3 = None, None
            ops.pool_add_tokens(body, skip, pos, tokens, b, t)
        del body, skip
        ops.token_fill(tokens, f32(vit.space_token.reshape(-1)), f32(vit.temporal_token.reshape(-1)), pos)
        layers = _pack_layers(vit, mirror)
        logits, ctxs, x_final, head = transformer_forward_train(vit, layers, tokens)
        saved = SimpleNamespace(layers=layers, ctxs=ctxs, x_final=x_final, head=head, ectx=ectx, amax3=amax3, b=b, t=t,
                                vit=vit, entry=entry)
        return logits, saved

    def backward(self, saved, dlogits: torch.Tensor, after_transformer=None) -> None:
        """Accumulates into the flat gradient buffer — call zero_grad() first unless accumulation is intended (every
        kernel on this path adds into its slot or reduces into a per-call scratch first, see istvt_bn_bwd).
        `after_transformer()` is called once every vit.* gradient is final (before the entry-flow backward)."""
        G = self.state.grad
        g = transformer_backward(saved.vit, saved.layers, saved.ctxs, saved.x_final, saved.head, dlogits, G)
        ops.token_bwd(g, G["vit.pos_embedding"][0], G["vit.space_token"].view(-1), G["vit.temporal_token"].view(-1))
        if after_transformer is not None:
            after_transformer()
        if saved.entry is not None:
            saved.entry.backward(saved.ectx, saved.amax3, g, G)

    def zero_grad(self) -> None:
        self.state.grads.zero_()

    def optimizer_step(self, world: int = 1) -> None:
        if self.replica:
            raise RuntimeError("an nn.DataParallel replica has no optimizer state: step the source module's optimizer")
        st = self.state
        st.step += 1
        ops.adamw_step(st.params, st.grads, st.exp_avg, st.exp_avg_sq, self.lr, self.betas, self.eps,
                       self.weight_decay, st.step, grad_scale=1.0 / world)
        self.model._engine = None            # the inference-mode packed weights are stale now

    def step(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One training iteration (train_CNN.py:513-533).  Returns the (local) loss."""
        import torch.distributed as dist
        self.zero_grad()
        logits, saved = self.forward_train(x)
        z = logits.view(-1)
        y = labels.to(z.device).float()
        loss = F.binary_cross_entropy_with_logits(z, y)                  # criterion, train_CNN.py:148,526
        dlogits = (torch.sigmoid(z) - y) / z.numel()                     # d(mean BCE)/dz
        world = dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            # Two buckets (SUM; AdamW scales by 1/world): the transformer's 353 MB are reduced on NCCL's stream while
            # the entry-flow backward (~1/6 of the backward) still runs; the entry flow's 4.4 MB follow at the end.
            st, pending = self.state, []
            vit_bucket, entry_bucket = st.grads[st.vit_offset:], st.grads[:st.vit_offset]
            self.backward(saved, dlogits, after_transformer=lambda: pending.append(
                dist.all_reduce(vit_bucket, group=self.pg, async_op=True)))
            if entry_bucket.numel():
                dist.all_reduce(entry_bucket, group=self.pg)
            for work in pending:
                work.wait()
        else:
            self.backward(saved, dlogits)
        self.optimizer_step(world)
        return loss


# ------------------------------------------------------------------------------------------------
# torch.autograd bridge: `outputs = model(image); loss.backward()` (train_CNN.py:517,532) in train mode
# ------------------------------------------------------------------------------------------------
class _ISTVTTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trainer, *params):
        logits, saved = trainer.forward_train(x)
        ctx.trainer, ctx.saved = trainer, saved
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        tr = ctx.trainer
        tr.zero_grad()                       # the flat buffer holds exactly this backward's gradients
        tr.backward(ctx.saved, dlogits.contiguous())
        ctx.saved = None
        return (None, None) + tuple(tr.state.grad[n].clone() for n in tr.state.names)


def bridge_trainer(model, dev: torch.device) -> "Trainer":
    """The Trainer behind `model(x)` in train mode, cached in the state a model shares with its `nn.DataParallel`
    replicas (network/vivit/vivit.py::_Shared): one for the module itself, one per device for replicas — replicas are
    rebuilt by every DataParallel.forward, so the cached Trainer is re-bound to the replica it is asked for."""
    sh = model._shared
    if getattr(model, "_is_replica", False):
        tr = sh.replica_trainers.get(dev.index)
        if tr is None:
            tr = Trainer(model)
            sh.replica_trainers[dev.index] = tr
        tr.model = model
        return tr
    if sh.trainer is None:
        sh.trainer = Trainer(model)
    return sh.trainer


def autograd_forward(model, x: torch.Tensor) -> torch.Tensor:
    """Training-mode `XceptionVidTr.forward`: the CUDA forward that keeps activations, differentiable through the
    hand-written backward.  Any torch optimizer over `model.parameters()` then works as in the reference, and so does
    `nn.DataParallel(model)` (train_CNN.py:185-186): on a replica the parameters handed to autograd are the broadcast
    copies, whose history reduces the per-replica gradients onto the source module."""
    tr = bridge_trainer(model, x.device)
    params = [p for _, p in on_path_named_parameters(model, tr.train_entry_flow)]
    return _ISTVTTrainFunction.apply(x, tr, *params)
