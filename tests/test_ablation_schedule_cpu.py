"""CPU: the HOST SCHEDULE of the ablation transformers (2023-tifs-istvt_b200/ablation.py) against the golden vectors of the
unmodified reference.

No CUDA kernel runs here.  Every C-ABI wrapper the schedule calls (`ops.gemm`, `ops.layernorm`, `ops.attn_spatial`,
`ops.attn_joint`, `ops.attn_temporal`, `ops.token_build`, `ops.gather_rows`, `ops.mean_rows`, `ops.pool_linear`,
`ops.head`) is replaced, for the duration of a test, by its torch fp32 definition — the same definitions
tests/kernel_checks.py holds each kernel to on the GPU.  What is pinned is therefore the part of the product that is NOT
a kernel: which op runs on which rows in which order, the weight packing, the token layouts, the class-row pruning of
the last layer (`cls_only`), the pooling variants, the q|k / v weight split of `TemporalOnlyAttention`.  The kernels
themselves are covered by `-m gpu`.  This is test scaffolding only: the product has no CPU path (tests/test_host.py).
"""
import importlib

import pytest
import torch

import torch_ops as tops
from helpers import (GOLDEN_ABLATION, ablation_oracle, build_ablation_block, build_ablation_model, fingerprint_check, pkg)


@pytest.fixture
def torch_ops(monkeypatch):
    """Substitute the C-ABI wrappers by their torch definitions (tests/torch_ops.py) for one test."""
    return tops.install(monkeypatch, pkg().ops)


@pytest.mark.parametrize("name", ["vivit_d2_b2", "vivit_mean_d2_b2", "vanilla_d2_b1"])
def test_model_schedules_match_reference_golden(torch_ops, name):
    A = ablation_oracle()
    ablation = importlib.import_module(pkg().__name__ + ".ablation")
    case = torch.load(GOLDEN_ABLATION, weights_only=False)["models"][name]
    model = build_ablation_model(case)
    x = A.make_features(case["batch"], 6)
    tol = 2e-6 * max(1.0, case["logits"].abs().max().item())
    # full schedule (intermediates tapped: no pruning)
    taps = {}
    full = ablation.features_forward(model, x, "fp32", taps)
    assert torch.allclose(full, case["logits"], rtol=0, atol=tol)
    n_full = len(torch_ops)
    # production schedule: the last layer runs on the class rows only after its attention
    del torch_ops[:]
    model.precision = "fp32"
    pruned = model(x)
    assert torch.allclose(pruned, case["logits"], rtol=0, atol=tol)
    assert torch_ops.count("gather_rows") >= 2                          # attention rows + residual rows of the class tokens
    assert len(torch_ops) <= n_full + 4
    # the long sequences go to the key-streaming kernel, the short ones to the all-keys-resident one
    if case["cls"] == "VanillaTr":
        assert "attn_joint" in torch_ops and "attn_spatial" not in torch_ops
    else:
        assert "attn_spatial" in torch_ops and "attn_joint" not in torch_ops
        assert ("mean_rows" in torch_ops) == (case.get("pool") == "mean")


@pytest.mark.parametrize("name", ["attention_n2167", "attention_n300", "temporal_only_t6"])
def test_block_schedules_match_reference_golden(torch_ops, name):
    A = ablation_oracle()
    case = torch.load(GOLDEN_ABLATION, weights_only=False)["blocks"][name]
    blk = build_ablation_block(case)
    blk.precision = "fp32"
    y = blk(A.make_tokens(case["batch"], case["n"]))
    fingerprint_check(f"schedule/{name}", y, case["out"], 2e-6)
    want_kernel = {"attention_n2167": "attn_joint", "attention_n300": "attn_spatial", "temporal_only_t6": "attn_temporal"}[name]
    assert want_kernel in torch_ops


def test_transformer_schedule_matches_oracle(torch_ops):
    A = ablation_oracle()
    torch.manual_seed(5)
    tr = pkg().Transformer(728, 2, 8, 64, 2912).eval()
    tr.precision = "fp32"
    sd = {"t." + k: v.clone() for k, v in tr.state_dict().items()}
    x = A.make_tokens(2, 40)
    with torch.no_grad():
        want = A.plain_transformer(sd, "t", x)
    assert torch.allclose(tr(x), want, rtol=0, atol=2e-6 * want.abs().max().item())
    # 2 layers x (2 LayerNorm, 4 GEMM, 1 attention) + the final LayerNorm
    assert torch_ops.count("gemm") == 8 and torch_ops.count("layernorm") == 5 and torch_ops.count("attn_spatial") == 2


def test_packed_weights_follow_parameter_updates(torch_ops):
    """The packed-weight cache is keyed on parameter storage and version counters: an in-place update (optimizer step,
    `load_state_dict`) must be picked up by the next forward."""
    A = ablation_oracle()
    torch.manual_seed(7)
    blk = pkg().Attention(728).eval()
    blk.precision = "fp32"
    x = A.make_tokens(1, 30)
    y0 = blk(x).clone()
    with torch.no_grad():
        blk.to_out[0].bias.add_(1.0)
    assert torch.allclose(blk(x), y0 + 1.0, atol=1e-5)
