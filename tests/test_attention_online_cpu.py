"""The arithmetic of the online-softmax spatial attention kernel (csrc/attn_spatial_pp.cuh, replaces
/root/reference/network/vivit/module.py:84-91), restated in torch (tests/torch_ops.py::attn_online_lazy_ref): visiting the
keys in steps with a reference that is raised lazily (only beyond 2^8) and rescaling the accumulated output must equal a
plain softmax(Q K^T) V — also when the raises actually happen.  The GPU checks `attn_spatial_bf16` / `attn_spatial_spiky`
compare the CUDA kernels with the same plain attention."""
import math

import torch

import torch_ops


def _plain(q, k, v, scale):
    s = torch.einsum("bid,bjd->bij", q, k) * scale
    a = torch.softmax(s, dim=-1)
    return torch.einsum("bij,bjd->bid", a, v), torch.logsumexp(s, dim=-1) / math.log(2.0)


def _qkv(items, n, seed, amp=1.0):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: (torch.randn(items, n, 64, generator=g) * amp).to(torch.bfloat16).float()
    return mk(), mk(), mk()


def test_online_equals_plain_without_raises():
    q, k, v = _qkv(3, 362, 0)
    cnt = []
    o, lse = torch_ops.attn_online_lazy_ref(q, k, v, 0.125, count=cnt)
    o_ref, lse_ref = _plain(q, k, v, 0.125)
    assert cnt[0] == 0                                   # random scores never exceed the first step's maximum by 2^8
    assert float((o - o_ref).abs().max()) <= 6e-3 * float(o_ref.abs().max())      # bf16 rounding of P only
    assert torch.allclose(lse, lse_ref, atol=1e-4)


def test_online_equals_plain_with_raises():
    # one key per row ~60 nats above the rest, in a LATER step: the reference must be raised and O / l rescaled
    for key in (100, 300, 361):
        q, k, v = _qkv(2, 362, key)
        q[:, :, 0] = 8.0
        k[:, :, 0] = 0.0
        k[:, key, 0] = 60.0                              # logit = 8 * 60 * 0.125 = 60 nats
        cnt = []
        o, lse = torch_ops.attn_online_lazy_ref(q, k, v, 0.125, count=cnt)
        o_ref, lse_ref = _plain(q, k, v, 0.125)
        assert cnt[0] >= 2 * 362                         # every row raised at least once
        assert float((o - o_ref).abs().max()) <= 6e-3 * float(o_ref.abs().max())
        assert torch.allclose(lse, lse_ref, atol=1e-3)


def test_lazy_threshold_keeps_p_in_range():
    # a maximum that grows by LESS than 2^8 per step is not chased: P may exceed 1 (up to 2^8) but stays finite in bf16
    q, k, v = _qkv(1, 256, 5, amp=2.0)
    cnt = []
    o, _ = torch_ops.attn_online_lazy_ref(q, k, v, 0.125, count=cnt)
    o_ref, _ = _plain(q, k, v, 0.125)
    assert torch.isfinite(o).all()
    assert float((o - o_ref).abs().max()) <= 8e-3 * float(o_ref.abs().max())
