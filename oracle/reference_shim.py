"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference model (Vill-Lab/2023-TIFS-ISTVT).

Only usable where the reference checkout exists (the build container: /root/reference).  Nothing under
tests marked `gpu`, `__graft_entry__.smoke()` or `bench.py` may depend on it at run time; it is used to
(a) pin oracle/istvt_oracle.py against the real modules and (b) generate tests/golden/*.pt
(oracle/make_golden.py).  It is never imported by the product package.

Recipe: the reference imports third-party packages that are not installed
and hard-loads /mnt/data/DFD/xception-b5690688.pth; we register empty stub modules for the former and
patch `return_pytorch04_xception(pretrained=False)` for the latter.  No reference source is modified.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ISTVT_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "efficientnet_pytorch", "efficientnet_pytorch.model", "vit_pytorch", "vit_pytorch.cvt", "vit_pytorch.cross_vit",
    "perceiver_pytorch", "rotary_embedding_torch", "attention_lib", "attention_lib.attention",
] + [f"attention_lib.attention.{n}" for n in (
    "OutlookAttention", "CoTAttention", "PolarizedSelfAttention", "CBAM", "S2Attention", "ShuffleAttention", "SGE",
    "PSA", "BAM")]


class _Anything(types.ModuleType):
    """Module whose every attribute is a dummy class (nothing on the ISTVT path touches them)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network", "vivit"))


_loaded = None


def load():
    """Returns the reference's `network.vivit.vivit` module (with `XceptionVidTr`, `DSTTr`)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, _Anything(name))
    # The product package also has a package called `network`; make sure the reference's wins here.
    for k in [k for k in sys.modules if k == "network" or k.startswith("network.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import network.models_copy as mc
            import network.xception as rx
            mc.return_pytorch04_xception = lambda pretrained=True: rx.return_pytorch04_xception(pretrained=False)
            import network.vivit.vivit as vv
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _loaded = vv
    return vv


def build_reference_model(seed: int = 0):
    """`torch.manual_seed(seed); XceptionVidTr()` of the reference, eval mode, fp32, CPU."""
    import torch
    vv = load()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = vv.XceptionVidTr()
    return m.eval()
