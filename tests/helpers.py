"""Shared test helpers: package loading, seeded weights, golden-fixture comparison."""
from __future__ import annotations

import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "2023-tifs-istvt_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden", "istvt_golden.pt")
GOLDEN_XCEPTION = os.path.join(ROOT, "tests", "golden", "xception_golden.pt")
GOLDEN_ABLATION = os.path.join(ROOT, "tests", "golden", "ablation_golden.pt")


def pkg():
    return importlib.import_module(PKG_NAME)


def oracle():
    from oracle import istvt_oracle
    return istvt_oracle


def make_input(batch: int, t: int, seed: int = 1234) -> torch.Tensor:
    """Same recipe as oracle/make_golden.py::make_input."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, t, 3, 300, 300, generator=g)
    x[1::2] = 2 * x[1::2] - 1
    return x


def build_model(case: dict):
    """Rebuild the weights of a golden case from its seeds (see oracle/make_golden.py)."""
    m = pkg()
    O = oracle()
    torch.manual_seed(case["seed"])
    model = m.XceptionVidTr(num_frames=6)
    if case["frames"] != 6:
        # make_golden.py replaces `vit` by a separately seeded DSTTr(19,1,1,T)
        model = _swap_vit(model, case["frames"], case["vit_seed"])
    if case["sensitised"]:
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        O.sensitise_(sd)
        model.load_state_dict(sd)
    return model.eval()


def _swap_vit(model, frames: int, vit_seed: int):
    m = pkg()
    torch.manual_seed(vit_seed)
    model.vit = m.DSTTr(19, 1, 1, frames)
    model.num_frames = frames
    return model


def fingerprint_check(name: str, got: torch.Tensor, want: dict, tol: float) -> float:
    """Compare a tensor with its golden fingerprint; returns the norm-wise relative error
    max|got - want| / max|want| over the sampled positions."""
    O = oracle()
    got = got.detach().float().cpu().contiguous()
    assert tuple(got.shape) == tuple(want["shape"]), f"{name}: shape {tuple(got.shape)} != {want['shape']}"
    flat = got.reshape(-1)
    idx = O.fingerprint_indices(flat.numel())
    err = (flat[idx] - want["samples"]).abs().max().item()
    rel = err / max(want["absmax"], 1e-30)
    assert rel <= tol, f"{name}: relative error {rel:.3e} > {tol:.1e} (absmax {want['absmax']:.3e})"
    return rel


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def make_frames(n: int, side: int, seed: int = 77) -> torch.Tensor:
    """Same recipe as oracle/make_golden_xception.py::make_frames."""
    g = torch.Generator().manual_seed(seed + side)
    x = torch.rand(n, 3, side, side, generator=g)
    x[1::2] = 2 * x[1::2] - 1
    return x


def build_xception(seed: int = 0):
    """The per-frame baseline `model_selection('xception', 2)` with the golden fixture's weights: seeded construction
    (the Xception tree is the first thing the reference's XceptionVidTr() draws, so the same seed gives the same
    values) + the deterministic `sensitise_xception_`."""
    torch.manual_seed(seed)
    model = pkg().model_selection("xception", num_out_classes=2, dropout=0.5)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    oracle().sensitise_xception_(sd, "model")
    model.load_state_dict(sd)
    return model.eval()


def ablation_oracle():
    from oracle import ablation_oracle as A
    return A


def build_ablation_model(case: dict):
    """Rebuild an ablation model of tests/golden/ablation_golden.pt from its seed with the B200 package's classes (same
    construction order as the reference, oracle/make_golden_ablation.py::build) + the deterministic sensitisation."""
    torch.manual_seed(case["seed"])
    model = getattr(pkg(), case["cls"])(19, 1, 1, 6, depth=case["depth"], pool=case.get("pool", "cls")).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ablation_oracle().sensitise_ablation_(sd)
    model.load_state_dict(sd)
    return model


def build_ablation_block(case: dict):
    torch.manual_seed(case["seed"])
    return getattr(pkg(), case["kind"])(728).eval()
