"""CPU: the HOST SCHEDULE of the ISTVT forward (2023-tifs-istvt_b200/engine.py: weight packing with BatchNorm folding, NHWC
entry flow, token build, 12 spatial-temporal blocks, last-block pruning, head; the uint8-frame stem; the variant models
behind the same entry flow) against the golden vectors of the unmodified reference and the CPU oracle.

No CUDA kernel runs here: every C-ABI wrapper is replaced by its torch fp32 definition (tests/torch_ops.py — the
definitions the GPU kernel checks hold the kernels to).  This pins the non-kernel half of the product in the
`-m "not gpu"` suite; the kernels are covered by `-m gpu`.  Test scaffolding only: the product has no CPU path.
"""
import pytest
import torch

import torch_ops as tops
from helpers import GOLDEN, ablation_oracle, build_model, make_input, oracle, pkg, rel_err


@pytest.fixture
def torch_ops(monkeypatch):
    return tops.install(monkeypatch, pkg().ops)


@pytest.mark.parametrize("name", ["default_init_b1", "sensitised_b2"])
def test_istvt_schedule_matches_reference_golden(torch_ops, name):
    case = torch.load(GOLDEN, weights_only=False)["cases"][name]
    model = build_model(case)
    model.precision = "fp32"
    x = make_input(case["batch"], case["frames"])
    # full schedule with attention maps (what the parity tests and the relevance pass read)
    logits, attn = model.engine().forward(model, x, precision="fp32", return_attention=True)
    assert rel_err(logits, case["logits"]) <= 1e-5
    assert len(attn) == 12 and tuple(attn[0][0].shape) == (case["batch"], 8, 362, 7, 7)
    per_layer = {k: torch_ops.count(k) / 12 for k in ("layernorm_diff", "attn_temporal", "attn_spatial")}
    assert per_layer == {"layernorm_diff": 1.0, "attn_temporal": 1.0, "attn_spatial": 1.0}
    assert torch_ops.count("dwconv3x3") == 6 and torch_ops.count("conv_stem") == 1 and torch_ops.count("conv3x3") == 1
    n_full = torch_ops.count("gemm")
    # production call: the last block is pruned to the rows that can reach token (0, 0)
    del torch_ops[:]
    pruned = model(x)
    assert rel_err(pruned, case["logits"]) <= 1e-5
    assert torch_ops.count("gemm") == n_full and torch_ops.count("gather_rows") == 3
    assert torch.equal(pruned > 0, case["logits"] > 0)


def test_uint8_frames_schedule_matches_oracle(torch_ops):
    """Decoded frames: the input normalisation folded into the stem weights (engine.fold_input_norm)."""
    O = oracle()
    torch.manual_seed(0)
    model = pkg().XceptionVidTr(precision="fp32").eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    O.sensitise_(sd)
    model.load_state_dict(sd)
    u8 = torch.randint(0, 256, (1, 6, 300, 300, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    with torch.no_grad():
        want = O.forward(sd, O.normalise_u8(u8))
    got = model(u8)
    assert "conv_stem_u8" in torch_ops and "conv_stem" not in torch_ops
    assert rel_err(got, want) <= 1e-4


@pytest.mark.parametrize("variant", ["vivit", "vanilla"])
def test_variant_models_behind_the_entry_flow(torch_ops, variant):
    A = ablation_oracle()
    torch.manual_seed(3)
    model = pkg().XceptionVidTr(variant=variant, precision="fp32").eval()
    # two layers are enough for the schedule: drop the rest of the (default depth 12) transformers
    for tr in ([model.vit.space_transformer, model.vit.temporal_transformer] if variant == "vivit" else [model.vit.transformer]):
        del tr.layers[2:]
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    oracle().sensitise_(sd)
    model.load_state_dict(sd)
    x = make_input(1, 6, seed=41)
    with torch.no_grad():
        want = A.clip_forward(sd, x, variant)
    got = model(x)
    assert rel_err(got, want) <= 1e-4 or (got - want).abs().max().item() <= 1e-5
    assert "pool_add" in torch_ops and "token_build" in torch_ops and "pool_add_tokens" not in torch_ops
