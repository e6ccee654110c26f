"""TEST INFRASTRUCTURE — generate tests/golden/xception_golden.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden_xception.py

The per-frame baseline of the reference is `model_selection('xception', num_out_classes=2, dropout=0.5)`
(network/models_copy.py:241-248 -> TransferModel -> network/xception.py Xception with a fresh Dropout + Linear(2048, 2)
head) — the very object `XceptionVidTr().xcep` holds (vivit.py:196).  It is evaluated on CPU in fp32, eval mode, and
fingerprints (oracle.istvt_oracle.fingerprint_indices) of block outputs, the bn4 features and the logits are stored.
Weights are rebuilt from the seed on the GPU box (same construction order as the reference, checked in
tests/test_oracle.py); `sensitise_xception_` is deterministic.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import istvt_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402
from oracle.make_golden import fp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "xception_golden.pt")


def make_frames(n: int, side: int, seed: int = 77) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + side)
    x = torch.rand(n, 3, side, side, generator=g)
    x[1::2] = 2 * x[1::2] - 1
    return x


def run_case(ref, x: torch.Tensor) -> dict:
    taps = {}
    hooks = []
    xm = ref.model
    for name in ("block3", "block4", "block11", "block12"):
        hooks.append(getattr(xm, name).register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.clone())))
    hooks.append(xm.bn4.register_forward_hook(lambda m, i, o: taps.__setitem__("features", o.clone())))
    with torch.no_grad():
        logits = ref(x)
    for h in hooks:
        h.remove()
    return {"logits": logits.detach().clone(), "taps": {k: fp(v) for k, v in taps.items()}}


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 8)
    ref = reference_shim.build_reference_model(seed=0).xcep.eval()      # TransferModel('xception')
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    O.sensitise_xception_(sd, "model")
    ref.load_state_dict(sd)
    golden = {"torch_version": torch.__version__, "seed": 0, "cases": {},
              "weights": {k: fp(sd[k]) for k in ("model.conv1.weight", "model.block7.rep.4.pointwise.weight",
                                                  "model.block12.skip.weight", "model.bn4.running_var",
                                                  "model.last_linear.1.weight")}}
    for name, n, side in (("b2_300", 2, 300), ("b1_299", 1, 299)):
        x = make_frames(n, side)
        c = run_case(ref, x)
        c.update({"n": n, "side": side})
        golden["cases"][name] = c
        print(name, c["logits"].tolist(), {k: round(v["absmax"], 3) for k, v in c["taps"].items()})
        # the oracle restatement must agree with the reference here and now
        want = O.xception_forward(sd, x, "model")
        print("   oracle vs reference:", (want - c["logits"]).abs().max().item())
    torch.save(golden, OUT)
    print("wrote", OUT, os.path.getsize(OUT) / 1e3, "KB")


if __name__ == "__main__":
    main()
