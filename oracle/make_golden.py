"""TEST INFRASTRUCTURE — generate tests/golden/istvt_golden.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

For each case the real `XceptionVidTr` (network/vivit/vivit.py:193-208) is evaluated on CPU in fp32 and a
compact fingerprint of the logits and of intermediate tensors is stored: shape, sum, max|x| and the
values at fixed pseudo-random positions (oracle.istvt_oracle.fingerprint_indices).  Intermediates are
observed without touching reference source: forward hooks on the reference's own sub-modules, and a
recording wrapper around `torch.Tensor.softmax` for the attention maps (module.py:88,202 call it as a
tensor method, which hooks cannot see).

Weights are never stored (437 MB): every case rebuilds them from a seed.  The product's module tree draws
the same initial values as the reference for the same seed (tests/test_oracle.py checks that here), and
`sensitise_` is deterministic, so the GPU box can regenerate them; a weight fingerprint in the fixture
guards against RNG drift.
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import istvt_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "istvt_golden.pt")


def fp(t: torch.Tensor) -> dict:
    t = t.detach().float().contiguous()
    flat = t.reshape(-1)
    idx = O.fingerprint_indices(flat.numel())
    return {"shape": tuple(t.shape), "sum": float(flat.double().sum()), "absmax": float(flat.abs().max()),
            "samples": flat[idx].clone()}


@contextlib.contextmanager
def record_softmax(store: list):
    orig = torch.Tensor.softmax

    def wrapped(self, *a, **k):
        out = orig(self, *a, **k)
        store.append(out)
        return out

    torch.Tensor.softmax = wrapped
    try:
        yield
    finally:
        torch.Tensor.softmax = orig


def make_input(batch: int, t: int, seed: int = 1234) -> torch.Tensor:
    """rand in [0,1) (test_time.py:7); odd clips use the 2x-1 normalisation (xception.py:12-13)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, t, 3, 300, 300, generator=g)
    x[1::2] = 2 * x[1::2] - 1
    return x


def run_case(ref, x: torch.Tensor, tap_layers=(0, 5, 11)) -> dict:
    taps = {}
    hooks = []
    xm = ref.xcep.model
    for name, mod in (("block1", xm.block1), ("block2", xm.block2), ("block3", xm.block3)):
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o)))
    hooks.append(xm.bn2.register_forward_hook(lambda m, i, o: taps.__setitem__("stem", torch.relu(o))))
    for li in tap_layers:
        for j, nm in enumerate(("temporal_out", "spatial_out", "ff_out")):
            hooks.append(ref.vit.transformer.layers[li][j].register_forward_hook(
                lambda m, i, o, key=f"layer{li}.{nm}": taps.__setitem__(key, o)))
    hooks.append(ref.vit.transformer.register_forward_hook(lambda m, i, o: taps.__setitem__("transformer_out", o)))
    soft = []
    with torch.no_grad(), record_softmax(soft):
        logits = ref(x)
    for h in hooks:
        h.remove()
    depth = len(ref.vit.transformer.layers)
    assert len(soft) == 2 * depth, len(soft)
    for li in tap_layers:
        taps[f"layer{li}.A_t"] = soft[2 * li]          # [b, h, hw, f, f]
        taps[f"layer{li}.A_s"] = soft[2 * li + 1]      # [b, h, f, hw, hw]
    out = {"logits": logits.detach().clone(), "taps": {k: fp(v) for k, v in taps.items()}}
    return out


def weight_fp(sd) -> dict:
    keys = ["xcep.model.conv1.weight", "xcep.model.block3.rep.4.pointwise.weight", "vit.pos_embedding",
            "vit.transformer.layers.11.2.fn.net.3.weight", "xcep.model.block2.rep.2.running_var",
            "vit.transformer.layers.3.1.norm.weight"]
    return {k: fp(sd[k]) for k in keys}


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 8)
    golden = {"torch_version": torch.__version__, "cases": {}}

    # case 1: default random init (BASELINE.json config 1: 1 clip, 6 frames, 300x300)
    ref = reference_shim.build_reference_model(seed=0)
    x = make_input(1, 6)
    c = run_case(ref, x)
    c.update({"seed": 0, "sensitised": False, "batch": 1, "frames": 6, "weights": weight_fp(ref.state_dict())})
    golden["cases"]["default_init_b1"] = c
    print("default_init_b1", c["logits"].flatten().tolist())

    # case 2: sensitised weights, 2 clips (one in [0,1), one in [-1,1))
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    O.sensitise_(sd)
    ref.load_state_dict(sd)
    x = make_input(2, 6)
    c = run_case(ref, x)
    c.update({"seed": 0, "sensitised": True, "batch": 2, "frames": 6, "weights": weight_fp(ref.state_dict())})
    golden["cases"]["sensitised_b2"] = c
    print("sensitised_b2", c["logits"].flatten().tolist())

    # case 3: long clip, 32 frames (BASELINE.json config 5) — reference XceptionVidTr with its `vit` replaced by
    # the reference's own DSTTr(19, 1, 1, 32) (vivit.py:201 hard-codes 6), seeded separately.
    vv = reference_shim.load()
    torch.manual_seed(1)
    ref.vit = vv.DSTTr(19, 1, 1, 32).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    O.sensitise_(sd)
    ref.load_state_dict(sd)
    x = make_input(1, 32)
    c = run_case(ref, x, tap_layers=(0, 11))
    c.update({"seed": 0, "vit_seed": 1, "sensitised": True, "batch": 1, "frames": 32,
              "weights": weight_fp(ref.state_dict())})
    golden["cases"]["sensitised_t32_b1"] = c
    print("sensitised_t32_b1", c["logits"].flatten().tolist())

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(golden, OUT)
    print("wrote", OUT, os.path.getsize(OUT) / 1e3, "KB")


if __name__ == "__main__":
    main()
