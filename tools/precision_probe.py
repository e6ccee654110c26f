"""Error attribution for the bf16 mode: run the golden cases with parts of the path switched to the fp32
validation kernels and print the logit error of each variant against the reference's golden logits.

    python tools/precision_probe.py            (on the GPU box; reads tests/golden/, no oracle needed)
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from helpers import GOLDEN, build_model, make_input, rel_err  # noqa: E402


def main() -> None:
    g = torch.load(GOLDEN, weights_only=False)
    for name in ("sensitised_t32_b1", "sensitised_b2", "default_init_b1"):
        case = g["cases"][name]
        model = build_model(case).cuda()
        x = make_input(case["batch"], case["frames"]).cuda()
        want = case["logits"]
        variants = {
            "bf16 (product)": dict(precision="bf16"),
            "bf16 + fp32 LN2 input": dict(precision="bf16", ln2_input_fp32=True),
            "entry fp32, transformer bf16": dict(precision="bf16", entry_precision="fp32"),
            "entry bf16, transformer fp32": dict(precision="fp32", entry_precision="bf16"),
            "fp32": dict(precision="fp32"),
        }
        for label, kw in variants.items():
            logits = model.engine().forward(model, x, **kw)
            torch.cuda.synchronize()
            print(f"{name:20s} {label:32s} rel_err={rel_err(logits, want):.3e} logits={[round(v, 5) for v in logits.flatten().tolist()]}"
                  f" want={[round(v, 5) for v in want.flatten().tolist()]}", flush=True)


if __name__ == "__main__":
    main()
