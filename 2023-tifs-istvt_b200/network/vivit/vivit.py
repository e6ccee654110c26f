"""ISTVT model = Xception entry flow + decomposed spatial-temporal transformer (B200-native forward).

API mirror of the reference's `network/vivit/vivit.py`: `STTransformer` (:85-101), `DSTTr` (:103-148),
`XceptionVidTr` (:193-208), and the ablation models `Transformer` (:10-25), `ViViT` (:29-81), `VanillaTr` (:150-191).  Constructor signatures, attribute paths (`xcep.model.*`,
`vit.transformer.layers[i][{0,1,2}].{norm,fn}`, `vit.mlp_head`) and therefore `state_dict` keys are the
reference's.  `forward` hands the clip to `engine.ISTVTEngine`, which runs the hand-written sm_100a
kernels through the C ABI; there is no torch-op or CPU fallback.
"""
from __future__ import annotations

import weakref

import torch
from torch import nn

from .module import Attention, FeedForward, PreNorm, SpatialOnlyAttention, TemporalResidualAttention


class Transformer(nn.Module):
    """depth x (PreNorm(Attention) + residual, PreNorm(FeedForward) + residual), final LayerNorm (reference
    vivit.py:10-25).  `forward(x: [b, n, dim] CUDA) -> [b, n, dim] fp32`, inference only."""

    def __init__(self, dim: int, depth: int, heads: int, dim_head: int, mlp_dim: int, dropout: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.norm = nn.LayerNorm(dim)
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]))
        self.precision = "bf16"

    def forward(self, x):
        from ...ablation import transformer_forward
        return transformer_forward(self, x, self.precision)


def _check_vit_args(pool, image_size, patch_size, in_channels, dim, dim_head, num_classes, patch_linear):
    if pool not in ("cls", "mean"):
        raise ValueError("pool type must be either cls (cls token) or mean (mean pooling)")
    if image_size % patch_size != 0:
        raise ValueError("Image dimensions must be divisible by the patch size.")
    if patch_size != 1 or dim_head != 64 or (in_channels != dim and not patch_linear):
        raise ValueError("the B200 path supports patch size 1, head dimension 64 and dim == channels")
    if num_classes != 1:
        raise ValueError("the B200 head kernel emits one logit (num_classes=1, vivit.py:201)")


class ViViT(nn.Module):
    """Factorised-encoder ablation (reference vivit.py:29-81): a `Transformer` over the 362 tokens of every frame, then a
    `Transformer` over the T+1 frame tokens of every clip.  Constructor signature and attribute names are the
    reference's; `forward(x: [b, t, 728, 19, 19] CUDA feature maps) -> logits [b, 1]` (inference only)."""

    def __init__(self, image_size: int, patch_size: int, num_classes: int, num_frames: int, dim: int = 728,
                 depth: int = 12, heads: int = 8, pool: str = "cls", in_channels: int = 728, dim_head: int = 64,
                 dropout: float = 0.0, emb_dropout: float = 0.0, scale_dim: int = 4):
        super().__init__()
        _check_vit_args(pool, image_size, patch_size, in_channels, dim, dim_head, num_classes, False)
        self.image_size, self.num_frames, self.dim, self.depth, self.heads = image_size, num_frames, dim, depth, heads
        num_patches = (image_size // patch_size) ** 2
        self.num_patches = num_patches
        # construction order of the reference (vivit.py:45-59), so that a seeded init draws the same values
        self.to_patch_embedding = nn.Sequential(nn.Identity())         # Rearrange only: no parameters (vivit.py:40-43)
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames, num_patches + 1, dim))
        self.space_token = nn.Parameter(torch.randn(1, 1, dim))
        self.space_transformer = Transformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.temporal_token = nn.Parameter(torch.randn(1, 1, dim))
        self.temporal_transformer = Transformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.dropout = nn.Dropout(emb_dropout)
        self.pool = pool
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self.precision = "bf16"

    def forward(self, x):
        from ...ablation import features_forward
        return features_forward(self, x, self.precision)


class VanillaTr(nn.Module):
    """Joint-attention ablation (reference vivit.py:150-191): per-patch Linear, then one `Transformer` over the
    T*361+1 tokens of a clip (2167 at T=6: the key-streaming attention kernel).
    `forward(x: [b, t, 728, 19, 19] CUDA feature maps) -> logits [b, 1]` (inference only)."""

    def __init__(self, image_size: int, patch_size: int, num_classes: int, num_frames: int, dim: int = 728,
                 depth: int = 12, heads: int = 8, pool: str = "cls", in_channels: int = 728, dim_head: int = 64,
                 dropout: float = 0.0, emb_dropout: float = 0.0, scale_dim: int = 4):
        super().__init__()
        _check_vit_args(pool, image_size, patch_size, in_channels, dim, dim_head, num_classes, True)
        if in_channels % 8 or dim % 8:
            raise ValueError("in_channels and dim must be multiples of 8")
        self.image_size, self.num_frames, self.dim, self.depth, self.heads = image_size, num_frames, dim, depth, heads
        num_patches = (image_size // patch_size) ** 2
        self.num_patches = num_patches
        patch_dim = in_channels * patch_size ** 2
        # index 1 carries the weights, like the reference's Sequential(Rearrange, Linear, Rearrange) (vivit.py:161-165)
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.Linear(patch_dim, dim), nn.Identity())
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames * num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.transformer = Transformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.dropout = nn.Dropout(emb_dropout)
        self.pool = pool
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self.precision = "bf16"

    def forward(self, x):
        from ...ablation import features_forward
        return features_forward(self, x, self.precision)


class STTransformer(nn.Module):
    def __init__(self, dim: int, depth: int, heads: int, dim_head: int, mlp_dim: int, dropout: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.norm = nn.LayerNorm(dim)
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, TemporalResidualAttention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, SpatialOnlyAttention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]))


class DSTTr(nn.Module):
    def __init__(self, image_size: int, patch_size: int, num_classes: int, num_frames: int, dim: int = 728,
                 depth: int = 12, heads: int = 8, pool: str = "cls", in_channels: int = 728, dim_head: int = 64,
                 dropout: float = 0.0, emb_dropout: float = 0.0, scale_dim: int = 4):
        super().__init__()
        if pool not in ("cls", "mean"):
            raise ValueError("pool type must be either cls (cls token) or mean (mean pooling)")
        if image_size % patch_size != 0:
            raise ValueError("Image dimensions must be divisible by the patch size.")
        if patch_size != 1 or in_channels != dim or dim_head != 64:
            raise ValueError("the B200 path supports the ISTVT configuration only (patch 1, dim == channels, head 64)")
        if num_classes != 1:
            raise ValueError("the B200 head kernel emits one logit (num_classes=1, vivit.py:201)")
        self.image_size, self.num_frames, self.dim, self.depth, self.heads = image_size, num_frames, dim, depth, heads
        num_patches = (image_size // patch_size) ** 2
        self.num_patches = num_patches
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames, num_patches + 1, dim))
        self.space_token = nn.Parameter(torch.randn(1, 1, dim))
        self.temporal_token = nn.Parameter(torch.randn(1, 1, dim))
        self.transformer = STTransformer(dim, depth, heads, dim_head, dim * scale_dim, dropout)
        self.dropout = nn.Dropout(emb_dropout)
        self.pool = pool
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))


class _Shared:
    """Derived state of one `XceptionVidTr` that its `nn.DataParallel` replicas must see too.  `replicate()` gives every
    replica a shallow copy of the module's `__dict__` (torch/nn/modules/module.py, `_replicate_for_data_parallel`) and
    rebuilds the replicas on EVERY forward, so anything cached on a replica is lost and anything keyed on the replica's
    parameter copies (fresh broadcast tensors each step) never hits.  This object is referenced, not copied: the engine
    (packed weights per device, keyed on the OWNER's parameter versions) and the per-device training state live here."""

    def __init__(self, owner):
        self.owner = weakref.ref(owner)
        self.engine = None
        self.trainer = None             # torch.autograd bridge of the owner (train.autograd_forward)
        self.replica_trainers = {}      # device index -> train.Trainer in replica mode

    def reset(self) -> None:
        self.engine = None
        self.trainer = None
        self.replica_trainers.clear()


class XceptionVidTr(nn.Module):
    """`XceptionVidTr()` as in the reference; `num_frames` / `precision` are additions with preserving defaults.

    forward(x: [B, T, 3, H, W] fp32 CUDA) -> logits [B, 1] fp32 (prediction = logit > 0, train_CNN.py:527).
    Inference also accepts decoded frames, uint8 [B, T, H, W, 3]: the normalisation `input_norm` = (mean, std) per
    channel on the [0, 1] scale — what the reference's (absent) `dataset.transform` does before the model,
    train_CNN.py:18-21,172-173; xception.py:12-13 documents 0.5 / 0.5 — is folded into the stem convolution.
    """
    input_norm = ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))

    VARIANTS = {"dsttr": DSTTr, "vivit": ViViT, "vanilla": VanillaTr}

    def __init__(self, num_frames: int = 6, precision: str = "bf16", variant: str = "dsttr"):
        super().__init__()
        from ..models import model_selection
        if variant not in self.VARIANTS:
            raise ValueError(f"variant must be one of {sorted(self.VARIANTS)}")
        self.xcep = model_selection(modelname="xception", num_out_classes=2, dropout=0.5, batch_size=1)
        # vivit.py:201 builds DSTTr(19, 1, 1, 6); the ablation transformers take the same arguments (vivit.py:30,151)
        self.vit = self.VARIANTS[variant](19, 1, 1, num_frames)
        self.variant = variant
        self.num_frames = num_frames
        self.precision = precision
        self._shared = _Shared(self)

    # `_engine` is kept as an attribute name (train.py resets it after an optimizer step); it lives in `_shared`
    @property
    def _engine(self):
        return self._shared.engine

    @_engine.setter
    def _engine(self, value):
        self._shared.engine = value

    def engine(self):
        from ...engine import ISTVTEngine
        sh = self._shared
        if sh.engine is None:
            owner = sh.owner()
            sh.engine = ISTVTEngine(owner if owner is not None else self)
        return sh.engine

    def __getstate__(self):              # torch.save(model) / copy.deepcopy: derived state is rebuilt, never serialised
        state = self.__dict__.copy()
        state.pop("_shared", None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__["_shared"] = _Shared(self)

    def forward(self, x: torch.Tensor, return_attention: bool = False):
        if not isinstance(self.vit, DSTTr):      # ablation transformer behind the same entry flow (inference only)
            if return_attention:
                raise ValueError("attention maps are an output of the ISTVT model (variant='dsttr') only")
            from ...ablation import clip_forward
            return clip_forward(self, x, self.precision)
        if self.training and torch.is_grad_enabled():
            # model.train() (train_CNN.py:226): BatchNorm batch statistics + activations kept for loss.backward()
            if return_attention:
                raise ValueError("attention maps are an inference-mode output")
            if not x.is_cuda:
                raise ValueError("ISTVT (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
            if x.dtype == torch.uint8:
                raise ValueError("uint8 clips are an inference-path input: normalise to fp32 [B, T, 3, H, W] for training")
            from ...train import autograd_forward
            return autograd_forward(self, x)
        if self.keep_relprop_input and not self.training:
            self.__dict__["_relprop_clips"] = x if x.shape[0] == 1 else None
        return self.engine().forward(self, x, precision=self.precision, return_attention=return_attention)

    keep_relprop_input = False      # True: eval-mode forwards remember their batch-1 clip for a following relprop()

    def relprop(self, cam=None, method: str = "transformer_attribution", is_ablation: bool = False,
                start_layer: int = 0, alpha: float = 1, clips=None):
        """Relevance maps (cam_s, cam_t) of the last decision — the call the reference's relevance generator makes on
        its model (visualize_rel.py:206,257; the model class itself, `tfe...ViT_LRP.VisionTransformer`, models.py:26,180,
        is absent from the reference tree).  See relevance.relprop."""
        if not isinstance(self.vit, DSTTr):
            raise NotImplementedError("relprop is built for the ISTVT model (variant='dsttr')")
        from ...relevance import relprop
        return relprop(self, cam, method=method, is_ablation=is_ablation, start_layer=start_layer, alpha=alpha, clips=clips)

    def save_packed_weights(self, path: str) -> str:
        """Persist the packed inference weights beside a checkpoint; see engine.save_pack."""
        from ...engine import save_pack
        return save_pack(self, path)

    def load_packed_weights(self, path: str) -> bool:
        """Install a persisted pack (False if it does not belong to the current weights); see engine.load_pack."""
        from ...engine import load_pack
        return load_pack(self, path)

    def _apply(self, fn, *args, **kwargs):  # .cuda() / .to(): drop the packed-weight cache and the flat train state
        self._shared.reset()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._shared.engine = None
        return super().load_state_dict(*args, **kwargs)
