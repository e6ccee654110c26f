"""CPU (`-m "not gpu"`): pin the oracle.

1. oracle/istvt_oracle.py vs the golden vectors that oracle/make_golden.py produced from the UNMODIFIED
   reference modules (logits, entry-flow taps, per-layer outputs, attention maps) — must be (near) bit exact;
2. where the reference checkout is present (build container), the oracle and the product's module tree are
   also checked against the live reference modules.
"""
import os

import pytest
import torch

from helpers import (GOLDEN, GOLDEN_ABLATION, GOLDEN_XCEPTION, ablation_oracle, build_ablation_block, build_ablation_model,
                     build_model, build_xception, fingerprint_check, make_frames, make_input, oracle)

TIGHT = 2e-6   # fp32 CPU vs fp32 CPU: the only freedom is thread-count dependent summation order


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _run_oracle(case):
    O = oracle()
    model = build_model(case)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = make_input(case["batch"], case["frames"])
    taps = {}
    with torch.no_grad():
        logits = O.forward(sd, x, taps)
    return model, logits, taps


@pytest.mark.parametrize("name", ["default_init_b1", "sensitised_b2"])
def test_oracle_matches_reference_golden(golden, name):
    case = golden["cases"][name]
    model, logits, taps = _run_oracle(case)
    for k, want in case["weights"].items():
        fingerprint_check(f"weights[{k}]", model.state_dict()[k], want, 0.0)
    assert torch.allclose(logits, case["logits"], rtol=0, atol=TIGHT * max(1.0, case["logits"].abs().max().item()))
    b, f = case["batch"], case["frames"] + 1
    for key, want in case["taps"].items():
        if key in ("stem", "block1", "block2", "block3"):
            got = taps[key]
        elif key.endswith(".A_t") or key.endswith(".A_s"):
            got = taps[key]
        elif key.endswith(".ff_out"):
            continue   # ff output before the residual is not a tap of the oracle; covered by transformer_out
        elif key.endswith(".temporal_out") or key.endswith(".spatial_out"):
            continue
        elif key == "transformer_out":
            continue
        else:
            raise AssertionError(f"unhandled golden tap {key}")
        fingerprint_check(f"{name}/{key}", got, want, TIGHT)


def test_oracle_layer_outputs_match_reference_hooks(golden):
    """temporal / spatial / ff sub-module outputs captured by forward hooks on the reference."""
    O = oracle()
    case = golden["cases"]["sensitised_b2"]
    model = build_model(case)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = make_input(case["batch"], case["frames"])
    with torch.no_grad():
        feats = O.entry_flow(sd, x.reshape(-1, 3, 300, 300)).reshape(2, 6, 728, 19, 19)
        h = O.build_tokens(sd, feats)
        for layer in range(12):
            lp = f"vit.transformer.layers.{layer}"
            y_t = O.temporal_attention(sd, f"{lp}.0.fn", O._ln(sd, f"{lp}.0.norm", h))
            y_s = O.spatial_attention(sd, f"{lp}.1.fn", O._ln(sd, f"{lp}.1.norm", y_t))
            h2 = y_s + h
            ff = O.feed_forward(sd, f"{lp}.2.fn", O._ln(sd, f"{lp}.2.norm", h2))
            for nm, got in (("temporal_out", y_t), ("spatial_out", y_s), ("ff_out", ff)):
                key = f"layer{layer}.{nm}"
                if key in case["taps"]:
                    fingerprint_check(key, got, case["taps"][key], TIGHT)
            h = ff + h2
        out = O._ln(sd, "vit.transformer.norm", h)
        fingerprint_check("transformer_out", out, case["taps"]["transformer_out"], TIGHT)


@pytest.mark.skipif(not os.path.isdir("/root/reference/network/vivit"), reason="reference checkout not present")
def test_against_live_reference():
    from oracle import reference_shim
    import importlib
    import sys
    O = oracle()
    ref = reference_shim.build_reference_model(seed=0)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    # the shim put the reference's `network` package in sys.modules; drop it before importing the product
    for k in [k for k in sys.modules if k == "network" or k.startswith("network.")]:
        del sys.modules[k]
    m = importlib.import_module("2023-tifs-istvt_b200")
    torch.manual_seed(0)
    mine = m.XceptionVidTr()
    msd = mine.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert msd[k].shape == sd[k].shape and torch.equal(msd[k], sd[k]), f"seeded init differs at {k}"
    O.sensitise_(sd)
    ref.load_state_dict(sd)
    x = make_input(1, 6, seed=5)
    with torch.no_grad():
        assert torch.equal(ref(x), O.forward(sd, x))


def test_oracle_training_step_against_reference_golden():
    """Oracle in train mode (BatchNorm batch statistics) + autograd + AdamW restatement vs the golden vectors that
    oracle/make_golden_train.py recorded from the UNMODIFIED reference (train_CNN.py:226,513-533)."""
    O = oracle()
    path = GOLDEN.replace("istvt_golden.pt", "istvt_golden_train.pt")
    g = torch.load(path, weights_only=False)
    model = build_model({"seed": g["seed"], "frames": g["frames"], "sensitised": g["sensitised"]})
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    before = {k: sd[k].clone() for k in g["params_after"]}
    x = make_input(g["batch"], g["frames"])
    loss, logits, grads = O.loss_and_grads(sd, x, torch.tensor(g["labels"]))
    assert abs(float(loss) - g["loss"]) <= 1e-6 * abs(g["loss"])
    assert torch.allclose(logits, g["logits"], rtol=0, atol=1e-6)
    assert sorted(grads) == sorted(g["grads"]) and len(grads) == 252
    assert len(g["params_without_grad"]) == 117
    def fp_err(t, w):
        flat = t.detach().float().reshape(-1)
        idx = O.fingerprint_indices(flat.numel(), count=256)
        return (flat[idx] - w["samples"]).abs().max().item() / max(w["absmax"], 1e-30)
    for k, gr in grads.items():
        assert fp_err(gr, g["grads"][k]) <= 1e-5, k
    for k, w in g["running_after"].items():
        assert fp_err(sd[k], w) <= 1e-6, k
    for k, gr in grads.items():
        p = before[k].clone()
        O.adamw_update(p, gr, torch.zeros_like(p), torch.zeros_like(p), 1, g["lr"], weight_decay=g["weight_decay"])
        assert fp_err(p, g["params_after"][k]) <= 1e-5, k


@pytest.mark.parametrize("name", ["b2_300", "b1_299"])
def test_xception_oracle_matches_reference_golden(name):
    """Per-frame Xception baseline: the oracle's xception_forward vs the golden vectors
    oracle/make_golden_xception.py recorded from the UNMODIFIED reference TransferModel('xception')."""
    O = oracle()
    g = torch.load(GOLDEN_XCEPTION, weights_only=False)
    case = g["cases"][name]
    model = build_xception(g["seed"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k, want in g["weights"].items():
        fingerprint_check(f"weights[{k}]", sd[k], want, 0.0)
    taps = {}
    with torch.no_grad():
        logits = O.xception_forward(sd, make_frames(case["n"], case["side"]), "model", taps)
    assert torch.allclose(logits, case["logits"], rtol=0, atol=TIGHT * max(1.0, case["logits"].abs().max().item()))
    for key, want in case["taps"].items():
        fingerprint_check(f"xception/{name}/{key}", taps[key], want, TIGHT)


# ------------------------------------------------------------------------------------------------
# ablation transformers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["vivit_d2_b2", "vivit_mean_d2_b2", "vanilla_d2_b1"])
def test_ablation_oracle_matches_reference_golden(name):
    """oracle/ablation_oracle.py vs the golden vectors oracle/make_golden_ablation.py recorded from the UNMODIFIED
    reference `ViViT` / `VanillaTr`; the weights are rebuilt from the seed with the B200 package's classes, so this
    also pins their construction order and state_dict keys."""
    A = ablation_oracle()
    case = torch.load(GOLDEN_ABLATION, weights_only=False)["models"][name]
    model = build_ablation_model(case)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k, want in case["weights"].items():
        fingerprint_check(f"weights[{k}]", sd[k], want, 0.0)
    taps = {}
    variant = "vivit" if case["cls"] == "ViViT" else "vanilla"
    with torch.no_grad():
        kw = {"pool": case.get("pool", "cls")} if variant == "vivit" else {}
        logits = A.FORWARDS[variant](sd, A.make_features(case["batch"], 6), "", taps, **kw)
    assert torch.allclose(logits, case["logits"], rtol=0, atol=TIGHT * max(1.0, case["logits"].abs().max().item()))
    rename = {"space_transformer_out": "space_out", "temporal_transformer_out": "temporal_out",
              "transformer_out": "transformer_out"}
    for key, want in case["taps"].items():
        fingerprint_check(f"ablation/{name}/{key}", taps[rename[key]], want, TIGHT)


@pytest.mark.parametrize("name", ["attention_n2167", "attention_n300", "temporal_only_t6"])
def test_ablation_block_oracle_matches_reference_golden(name):
    A = ablation_oracle()
    case = torch.load(GOLDEN_ABLATION, weights_only=False)["blocks"][name]
    blk = build_ablation_block(case)
    sd = {"a." + k: v for k, v in blk.state_dict().items()}
    fingerprint_check("weights[to_qkv]", sd["a.to_qkv.weight"], case["weights"]["to_qkv.weight"], 0.0)
    fn = A.joint_attention if case["kind"] == "Attention" else A.temporal_only_attention
    with torch.no_grad():
        y = fn(sd, "a", A.make_tokens(case["batch"], case["n"]))
    fingerprint_check(f"ablation/{name}", y, case["out"], TIGHT)


def test_ablation_state_dict_keys_match_live_reference():
    """Where the reference checkout exists: same state_dict keys / shapes as the reference's ViViT and VanillaTr."""
    from oracle import reference_shim
    if not reference_shim.available():
        pytest.skip("reference checkout not present")
    from helpers import pkg
    vv = reference_shim.load()
    for cls in ("ViViT", "VanillaTr"):
        ref = getattr(vv, cls)(19, 1, 1, 6, depth=1).state_dict()
        mine = getattr(pkg(), cls)(19, 1, 1, 6, depth=1).state_dict()
        assert list(ref.keys()) == list(mine.keys()), cls
        assert all(ref[k].shape == mine[k].shape for k in ref), cls
