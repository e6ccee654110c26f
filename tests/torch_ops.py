"""TEST SCAFFOLDING — torch fp32 definitions of the C-ABI wrappers in `<package>/ops.py`, used by the CPU schedule tests
(tests/test_ablation_schedule_cpu.py, tests/test_engine_schedule_cpu.py) to run the product's HOST SCHEDULE without a
GPU.  They are the same definitions tests/kernel_checks.py holds each CUDA kernel to on the B200.  Nothing in the
product imports this file; the product has no CPU path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gemm(a, w, bias=None, residual=None, act=0, out_dtype=None, out=None):
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


def layernorm(x, g, b, out_dtype, eps=1e-5, out=None):
    y = F.layer_norm(x.float(), (x.shape[-1],), g, b, eps).to(out_dtype)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def layernorm_diff(x, g, b, out_dtype, eps=1e-5, out=None):
    """LN + the self-subtract difference of module.py:192: frames 0 and 1 pass, frame f >= 2 becomes xn[f] - xn[f-1]."""
    xn = F.layer_norm(x.float(), (x.shape[-1],), g, b, eps)
    diff = torch.cat((xn[:, 0:2], xn[:, 2:] - xn[:, 1:-1]), dim=1)
    if out is not None:
        out[0].copy_(xn)
        out[1].copy_(diff)
        return out
    return xn.to(out_dtype), diff.to(out_dtype)


def attention(qkv, seqs, n, heads, scale):
    q, k, v = qkv.float().reshape(seqs, n, 3, heads, 64).permute(2, 0, 3, 1, 4)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    return (a @ v).permute(0, 2, 1, 3).reshape(seqs * n, heads * 64).to(qkv.dtype), a


def attn_joint(qkv, batch, tokens, heads, scale):
    return attention(qkv, batch, tokens, heads, scale)[0]


def attn_spatial(qkv, batch_frames, tokens, heads, scale, want_probs=False):
    o, a = attention(qkv, batch_frames, tokens, heads, scale)
    return o, (a if want_probs else None)


def attn_temporal(qk, v, b, f, p, heads, scale, want_probs=False):
    sp = lambda t: t.float().reshape(b, f, p, heads, 64).permute(0, 3, 2, 1, 4)      # b h p f d
    q, k, vv = sp(qk[:, :heads * 64]), sp(qk[:, heads * 64:]), sp(v)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    o = (a @ vv).permute(0, 3, 2, 1, 4).reshape(b * f * p, heads * 64).to(qk.dtype)
    return o, (a if want_probs else None)


def token_build(src, cls, pos, seqs, n, pos_period=1):
    dim = src.shape[-1]
    t = torch.cat((cls.reshape(1, 1, dim).expand(seqs, 1, dim), src.float().reshape(seqs, n, dim)), 1)
    if pos is not None:
        t = t + pos.reshape(pos_period, n + 1, dim)[torch.arange(seqs) % pos_period]
    return t.contiguous()


def gather_rows(src, n_outer, outer_stride, rows, row_stride, width):
    flat = src.reshape(-1)
    return torch.stack([flat[o * outer_stride + r * row_stride: o * outer_stride + r * row_stride + width]
                        for o in range(n_outer) for r in range(rows)])


def head(tokens, ng, nb, hg, hb, hw, hbias, eps=1e-5):
    x = tokens[:, 0, 0]
    x = F.layer_norm(x, (x.shape[-1],), ng, nb, eps)
    x = F.layer_norm(x, (x.shape[-1],), hg, hb, eps)
    return x @ hw.reshape(-1, 1) + hbias


def pool_linear(x, w, bias, relu=True):
    m = x.float().reshape(x.shape[0], -1, x.shape[-1])
    m = (F.relu(m) if relu else m).mean(1)
    return m @ w.t() + bias


def mean_rows(x, seqs, n):
    return x.reshape(seqs, n, -1).mean(1)


# ---- entry flow (NHWC activations) ----
def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def conv_stem(x, w, bias, out_dtype):
    return _nhwc(F.relu(F.conv2d(x, w, bias, stride=2)), out_dtype)


def conv_stem_u8(x, w, bias, out_dtype):
    return _nhwc(F.relu(F.conv2d(_nchw(x), w, bias, stride=2)), out_dtype)


def conv3x3(x, w, bias, act=1):
    y = F.conv2d(_nchw(x), w.float().permute(0, 3, 1, 2), bias)
    return _nhwc(F.relu(y) if act == 1 else y, x.dtype)


def dwconv3x3(x, w, relu_in):
    xi = _nchw(x)
    if relu_in:
        xi = F.relu(xi)
    c = xi.shape[1]
    return _nhwc(F.conv2d(xi, w.permute(2, 0, 1)[:, None].contiguous(), None, 1, 1, 1, groups=c), x.dtype)


def subsample2(x):
    return x[:, ::2, ::2, :].contiguous()


def pool_add(x, skip):
    return _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1), x.dtype) + skip.reshape(x.shape[0], (x.shape[1] - 1) // 2 + 1,
                                                                           (x.shape[2] - 1) // 2 + 1, x.shape[3])


def pool_add_tokens(x, skip, pos_emb, tokens, batch, t):
    y = pool_add(x, skip).float()                                      # [b*t, ho, wo, c]
    n, ho, wo, c = y.shape
    pos = pos_emb.reshape(t, ho * wo + 1, c)
    tokens[:, 1:, 1:, :] = y.reshape(batch, t, ho * wo, c) + pos[None, :, 1:, :]


def token_fill(tokens, space_token, temporal_token, pos_emb):
    b, f, p, d = tokens.shape
    pos = pos_emb.reshape(f - 1, p, d)
    tokens[:, 0] = temporal_token.reshape(1, 1, d)
    tokens[:, 1:, 0] = space_token.reshape(1, 1, d) + pos[None, :, 0]


ALL = dict(gemm=gemm, layernorm=layernorm, layernorm_diff=layernorm_diff, attn_joint=attn_joint, attn_spatial=attn_spatial,
           attn_temporal=attn_temporal, token_build=token_build, gather_rows=gather_rows, head=head, pool_linear=pool_linear,
           mean_rows=mean_rows, conv_stem=conv_stem, conv_stem_u8=conv_stem_u8, conv3x3=conv3x3, dwconv3x3=dwconv3x3,
           subsample2=subsample2, pool_add=pool_add, pool_add_tokens=pool_add_tokens, token_fill=token_fill)


def install(monkeypatch, ops_module):
    """Replace every wrapper in `ops_module` by its torch definition (recording the call order) and let CPU tensors pass
    the product's CUDA-only guards, for the duration of one test.  Returns the list the calls are recorded into."""
    calls = []

    def rec(name, fn):
        def wrapped(*a, **k):
            calls.append(name)
            return fn(*a, **k)
        return wrapped

    for name, fn in ALL.items():
        monkeypatch.setattr(ops_module, name, rec(name, fn))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    return calls
