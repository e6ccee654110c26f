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
import torch.nn.functional as F

from helpers import (GOLDEN_ABLATION, ablation_oracle, build_ablation_block, build_ablation_model, fingerprint_check, pkg)


def _gemm(a, w, bias=None, residual=None, act=0, out_dtype=None, out=None):
    y = F.linear(a.float().reshape(-1, a.shape[-1]), w.float(), bias)
    if act == 2:
        y = F.gelu(y)
    elif act == 1:
        y = F.relu(y)
    if residual is not None:
        y = y + residual.reshape(y.shape)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    if out_dtype is None:
        out_dtype = torch.float32 if (residual is not None or a.dtype == torch.float32) else a.dtype
    return y.to(out_dtype).reshape(*a.shape[:-1], w.shape[0])


def _layernorm(x, g, b, out_dtype, eps=1e-5, out=None):
    return F.layer_norm(x.float(), (x.shape[-1],), g, b, eps).to(out_dtype)


def _attn(qkv, seqs, n, heads, scale):
    q, k, v = qkv.float().reshape(seqs, n, 3, heads, 64).permute(2, 0, 3, 1, 4)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    return (a @ v).permute(0, 2, 1, 3).reshape(seqs * n, heads * 64).to(qkv.dtype)


def _attn_temporal(qk, v, b, f, p, heads, scale, want_probs=False):
    sp = lambda t: t.float().reshape(b, f, p, heads, 64).permute(0, 3, 2, 1, 4)
    q, k, vv = sp(qk[:, :512]), sp(qk[:, 512:]), sp(v)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    return (a @ vv).permute(0, 3, 2, 1, 4).reshape(b * f * p, 512).to(qk.dtype), None


def _token_build(src, cls, pos, seqs, n, pos_period=1):
    dim = src.shape[-1]
    t = torch.cat((cls.reshape(1, 1, dim).expand(seqs, 1, dim), src.float().reshape(seqs, n, dim)), 1)
    if pos is not None:
        t = t + pos.reshape(pos_period, n + 1, dim)[torch.arange(seqs) % pos_period]
    return t.contiguous()


def _gather_rows(src, n_outer, outer_stride, rows, row_stride, width):
    flat = src.reshape(-1)
    return torch.stack([flat[o * outer_stride + r * row_stride: o * outer_stride + r * row_stride + width]
                        for o in range(n_outer) for r in range(rows)])


def _head(tokens, ng, nb, hg, hb, hw, hbias, eps=1e-5):
    x = tokens[:, 0, 0]
    x = F.layer_norm(x, (x.shape[-1],), ng, nb, eps)
    x = F.layer_norm(x, (x.shape[-1],), hg, hb, eps)
    return x @ hw.reshape(-1, 1) + hbias


def _pool_linear(x, w, bias, relu=True):
    m = x.float().reshape(x.shape[0], -1, x.shape[-1])
    m = (F.relu(m) if relu else m).mean(1)
    return m @ w.t() + bias


@pytest.fixture
def torch_ops(monkeypatch):
    """Substitute the C-ABI wrappers by their torch definitions and let CPU tensors pass the CUDA-only guards."""
    ops = pkg().ops
    calls = []

    def rec(name, fn):
        def wrapped(*a, **k):
            calls.append(name)
            return fn(*a, **k)
        return wrapped

    for name, fn in dict(gemm=_gemm, layernorm=_layernorm, attn_joint=_attn, attn_temporal=_attn_temporal,
                         attn_spatial=lambda qkv, bf, n, heads, scale, want_probs=False: (_attn(qkv, bf, n, heads, scale), None),
                         token_build=_token_build, gather_rows=_gather_rows, head=_head, pool_linear=_pool_linear,
                         mean_rows=lambda x, seqs, n: x.reshape(seqs, n, -1).mean(1)).items():
        monkeypatch.setattr(ops, name, rec(name, fn))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    return calls


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
