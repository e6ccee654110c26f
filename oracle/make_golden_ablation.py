"""TEST INFRASTRUCTURE — generate tests/golden/ablation_golden.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden_ablation.py

The ablation models of the reference (`ViViT` vivit.py:29-81, `VanillaTr` vivit.py:150-191, built on `Transformer`
vivit.py:10-25 / `Attention` module.py:36-64) and the stand-alone `TemporalOnlyAttention` (module.py:145-172) are
constructed from a seed with the reference's own classes (oracle/reference_shim.py; no source change), sensitised
(oracle.ablation_oracle.sensitise_ablation_, deterministic) and evaluated on CPU in fp32, eval mode, on seeded
stand-ins for the block-3 feature maps.  Logits, fingerprints of the transformer outputs and of a few weights are
stored; the GPU box rebuilds the weights from the same seeds with the B200 package's classes (same construction order).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ablation_oracle as A  # noqa: E402
from oracle import reference_shim  # noqa: E402
from oracle.make_golden import fp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ablation_golden.pt")

# name -> (class, depth, batch, seed[, pool])
MODEL_CASES = {
    "vivit_d2_b2": ("ViViT", 2, 2, 11),
    "vivit_mean_d2_b2": ("ViViT", 2, 2, 15, "mean"),
    "vivit_d12_b1": ("ViViT", 12, 1, 12),
    "vanilla_d2_b1": ("VanillaTr", 2, 1, 13),
    "vanilla_d12_b1": ("VanillaTr", 12, 1, 14),
}
WEIGHT_KEYS = {
    "ViViT": ("pos_embedding", "space_transformer.layers.1.0.fn.to_qkv.weight",
              "temporal_transformer.layers.0.1.fn.net.3.weight", "temporal_transformer.norm.weight", "mlp_head.1.weight"),
    "VanillaTr": ("pos_embedding", "to_patch_embedding.1.weight", "transformer.layers.1.0.fn.to_out.0.weight",
                  "transformer.layers.0.1.norm.bias", "mlp_head.1.weight"),
}


def build(vv, cls_name: str, depth: int, seed: int, pool: str = "cls"):
    torch.manual_seed(seed)
    m = getattr(vv, cls_name)(19, 1, 1, 6, depth=depth, pool=pool).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    A.sensitise_ablation_(sd)
    m.load_state_dict(sd)
    return m, sd


def run_model_case(vv, cls_name: str, depth: int, batch: int, seed: int, pool: str = "cls") -> dict:
    m, sd = build(vv, cls_name, depth, seed, pool)
    x = A.make_features(batch, 6)
    taps = {}
    hooks = []
    for name in ("space_transformer", "temporal_transformer", "transformer"):
        if hasattr(m, name):
            hooks.append(getattr(m, name).register_forward_hook(
                lambda mod, i, o, name=name: taps.__setitem__(name + "_out", o.detach().clone())))
    with torch.no_grad():
        logits = m(x.clone())
    for h in hooks:
        h.remove()
    variant = "vivit" if cls_name == "ViViT" else "vanilla"
    with torch.no_grad():
        want = A.FORWARDS[variant](sd, x, pool=pool) if variant == "vivit" else A.FORWARDS[variant](sd, x)
    print(f"   oracle vs reference: {(want - logits).abs().max().item():.3e}")
    return {"cls": cls_name, "depth": depth, "batch": batch, "seed": seed, "pool": pool, "logits": logits.detach().clone(),
            "taps": {k: fp(v) for k, v in taps.items()}, "weights": {k: fp(sd[k]) for k in WEIGHT_KEYS[cls_name]}}


def run_block_cases(vv) -> dict:
    mod = sys.modules["network.vivit.module"]
    out = {}
    # joint attention over a VanillaTr-length sequence (2167 tokens) and over a ragged short one
    for name, n, b, seed in (("attention_n2167", 2167, 1, 21), ("attention_n300", 300, 2, 22)):
        torch.manual_seed(seed)
        blk = mod.Attention(728).eval()
        x = A.make_tokens(b, n)
        with torch.no_grad():
            y = blk(x)
            want = A.joint_attention({"a." + k: v for k, v in blk.state_dict().items()}, "a", x)
        print(f"{name}: oracle vs reference {(want - y).abs().max().item():.3e}")
        out[name] = {"kind": "Attention", "n": n, "batch": b, "seed": seed, "out": fp(y),
                     "weights": {"to_qkv.weight": fp(blk.state_dict()["to_qkv.weight"])}}
    torch.manual_seed(23)
    blk = mod.TemporalOnlyAttention(728).eval()
    x = A.make_tokens(2, 6 * 362)
    with torch.no_grad():
        y = blk(x)
        want = A.temporal_only_attention({"a." + k: v for k, v in blk.state_dict().items()}, "a", x)
    print(f"temporal_only: oracle vs reference {(want - y).abs().max().item():.3e}")
    out["temporal_only_t6"] = {"kind": "TemporalOnlyAttention", "n": 6 * 362, "batch": 2, "seed": 23, "out": fp(y),
                               "weights": {"to_qkv.weight": fp(blk.state_dict()["to_qkv.weight"])}}
    return out


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 8)
    vv = reference_shim.load()
    golden = {"torch_version": torch.__version__, "models": {}, "blocks": run_block_cases(vv)}
    for name, spec in MODEL_CASES.items():
        print(name)
        c = run_model_case(vv, *spec)
        golden["models"][name] = c
        print("   logits", c["logits"].flatten().tolist(), {k: round(v["absmax"], 3) for k, v in c["taps"].items()})
    torch.save(golden, OUT)
    print("wrote", OUT, os.path.getsize(OUT) / 1e3, "KB")


if __name__ == "__main__":
    main()
