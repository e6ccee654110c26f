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


def conv3x3(x, w, bias, act=1, kernel=None):
    y = F.conv2d(_nchw(x), w.float().permute(0, 3, 1, 2), bias)
    return _nhwc(F.relu(y) if act == 1 else y, x.dtype)


def dwconv3x3(x, w, relu_in):
    xi = _nchw(x)
    if relu_in:
        xi = F.relu(xi)
    c = xi.shape[1]
    return _nhwc(F.conv2d(xi, w.permute(2, 0, 1)[:, None].contiguous(), None, 1, 1, 1, groups=c), x.dtype)


def sepconv_fused(x, dw, pw, bias, relu_in, act=0):
    d = dwconv3x3(x, dw, relu_in)
    n, h, w, _ = x.shape
    return gemm(d, pw, bias=bias, act=act).reshape(n, h, w, -1)


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


def gemm_act_dual(a, w, bias, act=2):
    pre = F.linear(a.float().reshape(-1, a.shape[-1]), w.float(), bias)
    out = F.gelu(pre) if act == 2 else (F.relu(pre) if act == 1 else pre)
    shape = (*a.shape[:-1], w.shape[0])
    return out.to(a.dtype).reshape(shape), pre.to(a.dtype).reshape(shape)


def gemm_dgelu(a, w, pre):
    y = F.linear(a.float().reshape(-1, a.shape[-1]), w.float())
    xf = pre.float().reshape(y.shape)
    cdf = 0.5 * (1 + torch.erf(xf * 0.7071067811865476))
    pdf = torch.exp(-0.5 * xf * xf) * 0.3989422804014327
    return (y * (cdf + xf * pdf)).to(a.dtype)


def gemm_rowstats(a, w, bias, row_stats):
    y = gemm(a, w, bias=bias)
    yr = y.float().reshape(-1, y.shape[-1])
    n = yr.shape[1]
    for g in range((n + 63) // 64):
        blk = yr[:, 64 * g: 64 * g + 64]
        row_stats.view(yr.shape[0], -1, 2)[:, g, 0] = blk.sum(1)
        row_stats.view(yr.shape[0], -1, 2)[:, g, 1] = ((blk - blk.mean(1, keepdim=True)) ** 2).sum(1)
    return y


def ln_stats_finalize(row_stats, dim, eps=1e-5):
    st = row_stats.reshape(-1, (dim + 63) // 64, 2)
    ng = torch.tensor([min(64, dim - 64 * g) for g in range(st.shape[1])], dtype=torch.float32)
    mu = st[:, :, 0].sum(1) / dim
    m2 = (st[:, :, 1] + ng * (st[:, :, 0] / ng - mu[:, None]) ** 2).sum(1)      # Chan's parallel-variance combination
    return torch.stack((mu, torch.rsqrt(m2 / dim + eps)), 1)


def gemm_lnfold(a, w_folded, mu_rstd, w_rowsum, shift):
    k = a.shape[-1]
    mu, rstd = mu_rstd.reshape(-1, 2).unbind(1)
    acc = a.float().reshape(-1, k) @ w_folded.float().t()
    return (rstd[:, None] * (acc - mu[:, None] * w_rowsum[None, :]) + shift[None, :]).to(a.dtype).reshape(*a.shape[:-1], -1)


# ---- training step: forward variants that keep what the backward needs, and the backward of every op ----
def conv_stem_raw(x, w):
    return _nhwc(F.conv2d(x, w, None, stride=2), x.dtype)


class _BNState:
    pass


def batchnorm_train(x, bn, relu, update_running=True):
    """nn.BatchNorm2d in train mode on NHWC rows: batch statistics, running statistics updated with the UNBIASED variance."""
    c = x.shape[-1]
    xf = x.float().reshape(-1, c)
    m = xf.shape[0]
    mean, var = xf.mean(0), xf.var(0, unbiased=False)
    st = _BNState()
    st.m, st.c, st.mean, st.rstd = m, c, mean, torch.rsqrt(var + bn.eps)
    st.scale = bn.weight.detach().float() * st.rstd
    st.shift = bn.bias.detach().float() - mean * st.scale
    if update_running:
        with torch.no_grad():
            bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
            bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * m / (m - 1))
            bn.num_batches_tracked += 1
    y = xf * st.scale + st.shift
    return (F.relu(y) if relu else y).reshape(x.shape).to(x.dtype), st


def batchnorm_bwd(dy, x, st, dgamma, dbeta, relu):
    c = st.c
    xf, g = x.float().reshape(-1, c), dy.float().reshape(-1, c)
    if relu:
        g = g * ((xf * st.scale + st.shift) > 0)
    xhat = (xf - st.mean) * st.rstd
    sb, sg = g.sum(0), (g * xhat).sum(0)
    dbeta += sb
    dgamma += sg
    return (st.scale * (g - sb / st.m - xhat * sg / st.m)).reshape(x.shape).to(x.dtype)


def pool_add_idx(x, skip, tokens=None, pos_emb=None, t_frames=1):
    """max-pool 3x3 s2 p1 + skip; the arg-max is kept as torch's flat window index (private to pool_bwd below)."""
    y, idx = F.max_pool2d(_nchw(x), 3, 2, 1, return_indices=True)
    y = _nhwc(y, torch.float32) + skip.float().reshape(x.shape[0], y.shape[2], y.shape[3], x.shape[3])
    if tokens is None:
        return y.to(x.dtype), idx
    n, ho, wo, c = y.shape
    pos = pos_emb.reshape(t_frames, ho * wo + 1, c)
    tokens[:, 1:, 1:, :] = y.reshape(n // t_frames, t_frames, ho * wo, c) + pos[None, :, 1:, :]
    return None, idx


def pool_bwd(dy, amax, h, wd):
    g = _nchw(dy)                                  # overlapping 3x3 / stride-2 windows can share an arg-max: accumulate
    dx = torch.zeros(g.shape[0], g.shape[1], h * wd)
    dx.scatter_add_(2, amax.flatten(2), g.flatten(2))
    return _nhwc(dx.view(g.shape[0], g.shape[1], h, wd), dy.dtype)


def token_grad_gather(g):
    b, f, p, c = g.shape
    side = int(round((p - 1) ** 0.5))
    return g[:, 1:, 1:, :].reshape(b * (f - 1), side, side, c).contiguous()


def token_bwd(g, d_pos, d_space, d_temporal):
    d_pos += g[:, 1:].sum(0)                       # pos_embedding is added to every row of the real frames (vivit.py:138)
    d_space += g[:, 1:, 0].sum((0, 1))
    d_temporal += g[:, 0].sum((0, 1))


def dwconv3x3_wgrad(x, dy, dw, relu_in):
    xi = F.pad(F.relu(_nchw(x)) if relu_in else _nchw(x), (1, 1, 1, 1))
    g = _nchw(dy)
    h, w = g.shape[2:]
    for ky in range(3):
        for kx in range(3):
            dw[ky, kx] += (xi[:, :, ky:ky + h, kx:kx + w] * g).sum((0, 2, 3))


def block_input_grad(d_main, x_in, d_skip, relu_in):
    dx = d_main.float() * (x_in > 0) if relu_in else d_main.float().clone()
    dx[:, ::2, ::2, :] += d_skip.float()
    return dx.to(d_main.dtype)


def transpose(x, colsum=None, aligned=False):
    if colsum is not None:
        colsum += x.float().sum(0)
    m = x.shape[0]
    if aligned:        # the product returns the [:, :m] view of a row-pitched buffer
        return x.t().contiguous()
    return F.pad(x.t(), (0, (m + 7) // 8 * 8 - m)).contiguous()


def im2col_t(x):
    n, h, w, cin = x.shape
    cols = torch.stack([x[:, ky:ky + h - 2, kx:kx + w - 2, :] for ky in range(3) for kx in range(3)])   # [9, n, ho, wo, cin]
    out = cols.permute(0, 4, 1, 2, 3).reshape(9 * cin, -1)
    return F.pad(out, (0, (out.shape[1] + 7) // 8 * 8 - out.shape[1])).contiguous()


def im2col_t_stem(x):
    n, _, h, w = x.shape
    ho, wo = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    rows = [x[:, c, ky:ky + 2 * ho - 1:2, kx:kx + 2 * wo - 1:2].reshape(-1) for c in range(3) for ky in range(3) for kx in range(3)]
    out = torch.cat((torch.stack(rows), torch.zeros(5, rows[0].numel())))
    return F.pad(out, (0, (out.shape[1] + 7) // 8 * 8 - out.shape[1])).contiguous()


def gemm_wgrad(dyt, xt, rows, dw):
    dw += (dyt[:, :rows].float() @ xt[:, :rows].float().t()).reshape(dw.shape)


def colsum(x, out):
    out += x.float().sum(0)


def wgrad(dy, x, dw, bias_grad=None):
    dw += (dy.float().t() @ x.float()).reshape(dw.shape)
    if bias_grad is not None:
        bias_grad += dy.float().sum(0)


def gelu(x):
    return F.gelu(x.float()).to(x.dtype)


def gelu_bwd(dy, x, colsum=None):
    xf = x.float()
    cdf = 0.5 * (1 + torch.erf(xf * 0.7071067811865476))
    pdf = torch.exp(-0.5 * xf * xf) * 0.3989422804014327
    dx = dy.float() * (cdf + xf * pdf)
    if colsum is not None:
        colsum += dx.reshape(-1, x.shape[-1]).sum(0)
    return dx.to(x.dtype)


def cast_bf16(x, out=None):
    if out is None:
        return x.clone()
    out.copy_(x)
    return out


def layernorm_bwd(dy, x, gamma, dgamma, dbeta, g_accum=None, g_bf16=None, dy2=None, frames=0, tokens_per_frame=0, eps=1e-5,
                  out_colsum=None):
    d = x.shape[-1]
    xf = x.float().reshape(-1, d)
    g = dy.float().reshape(-1, d).clone()
    if dy2 is not None:      # backward of the self-subtract (module.py:192) folded in: diff[f] = xn[f] - xn[f-1] for f >= 2
        gv, d2 = g.view(-1, frames, tokens_per_frame, d), dy2.float().reshape(-1, frames, tokens_per_frame, d)
        gv += d2
        gv[:, 1:frames - 1] -= d2[:, 2:frames]
    mean = xf.mean(-1, keepdim=True)
    rstd = torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (xf - mean) * rstd
    dgamma += (g * xhat).sum(0)
    dbeta += g.sum(0)
    gg = g * gamma
    dx = rstd * (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True))
    if g_accum is None:
        if out_colsum is not None:
            out_colsum += dx.sum(0)
        return dx.reshape(x.shape).to(dy.dtype)
    g_accum += dx.reshape(g_accum.shape)
    if out_colsum is not None:
        out_colsum += g_accum.reshape(-1, d).sum(0)
    if g_bf16 is not None:
        g_bf16.copy_(g_accum.reshape(g_bf16.shape))
    return None


def head_bwd(tokens, dlogits, ng, nb, hg, hb, hw, g, d_ng, d_nb, d_hg, d_hb, d_hw, d_hbias, eps=1e-5):
    with torch.enable_grad():
        leaves = [t.detach().clone().requires_grad_(True) for t in (tokens[:, 0, 0], ng, nb, hg, hb, hw)]
        x, a, b, c, e, w = leaves
        y = F.layer_norm(F.layer_norm(x, (x.shape[-1],), a, b, eps), (x.shape[-1],), c, e, eps)
        grads = torch.autograd.grad(y @ w.reshape(-1, 1), leaves, dlogits.reshape(-1, 1))
    g[:, 0, 0] += grads[0]
    for dst, src in zip((d_ng, d_nb, d_hg, d_hb, d_hw), grads[1:]):
        dst += src.reshape(dst.shape)
    d_hbias += dlogits.sum()


def attn_spatial_lse(qkv, batch_frames, tokens, heads, scale):
    return attention(qkv, batch_frames, tokens, heads, scale)[0], torch.zeros(batch_frames, heads, tokens)


def _attn_bwd(q, k, v, do, scale):
    """q, k, v, do: [..., n, 64] -> (dq, dk, dv, a, da)."""
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    da = do @ v.transpose(-1, -2)
    ds = a * (da - (a * da).sum(-1, keepdim=True)) * scale
    return ds @ k, ds.transpose(-1, -2) @ q, a.transpose(-1, -2) @ do, a, da


def attn_spatial_bwd(qkv, o, dout, lse, batch_frames, tokens, heads, scale, scratch=None, cam=None):
    q, k, v = qkv.float().reshape(batch_frames, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    do = dout.float().reshape(batch_frames, tokens, heads, 64).permute(0, 2, 1, 3)
    dq, dk, dv, a, da = _attn_bwd(q, k, v, do, scale)
    if cam is not None:
        cam += (a * da).clamp(min=0).mean(1).reshape(cam.shape)
    return torch.stack((dq, dk, dv)).permute(1, 3, 0, 2, 4).reshape(qkv.shape).to(qkv.dtype)


def attn_temporal_bwd(qk, v, dout, b, f, p, heads, scale, cam=None):
    sp = lambda t: t.float().reshape(b, f, p, heads, 64).permute(0, 3, 2, 1, 4)      # b h p f d
    inner = heads * 64
    dq, dk, dv, a, da = _attn_bwd(sp(qk[:, :inner]), sp(qk[:, inner:]), sp(v), sp(dout), scale)
    if cam is not None:
        cam += (a * da).clamp(min=0).mean(1).reshape(cam.shape)
    back = lambda t: t.permute(0, 3, 2, 1, 4).reshape(b * f * p, inner)
    return torch.cat((back(dq), back(dk)), 1).to(qk.dtype), back(dv).to(v.dtype)


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step=1, grad_scale=1.0):
    """torch.optim.AdamW (decoupled weight decay), one step on flat buffers."""
    g = grads * grad_scale
    params.mul_(1 - lr * weight_decay)
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    denom = (exp_avg_sq / (1 - betas[1] ** step)).sqrt_().add_(eps)
    params.addcdiv_(exp_avg / (1 - betas[0] ** step), denom, value=-lr)


ALL = dict(gemm=gemm, layernorm=layernorm, layernorm_diff=layernorm_diff, attn_joint=attn_joint, attn_spatial=attn_spatial,
           attn_temporal=attn_temporal, token_build=token_build, gather_rows=gather_rows, head=head, pool_linear=pool_linear,
           mean_rows=mean_rows, conv_stem=conv_stem, conv_stem_u8=conv_stem_u8, conv3x3=conv3x3, dwconv3x3=dwconv3x3,
           subsample2=subsample2, pool_add=pool_add, pool_add_tokens=pool_add_tokens, token_fill=token_fill,
           gemm_rowstats=gemm_rowstats, ln_stats_finalize=ln_stats_finalize, gemm_lnfold=gemm_lnfold, conv_stem_raw=conv_stem_raw, batchnorm_train=batchnorm_train, batchnorm_bwd=batchnorm_bwd, pool_add_idx=pool_add_idx,
           pool_bwd=pool_bwd, token_grad_gather=token_grad_gather, token_bwd=token_bwd, dwconv3x3_wgrad=dwconv3x3_wgrad,
           block_input_grad=block_input_grad, transpose=transpose, im2col_t=im2col_t, im2col_t_stem=im2col_t_stem,
           gemm_wgrad=gemm_wgrad, colsum=colsum, wgrad=wgrad, gelu=gelu, gelu_bwd=gelu_bwd, cast_bf16=cast_bf16,
           layernorm_bwd=layernorm_bwd, head_bwd=head_bwd, attn_spatial_lse=attn_spatial_lse, attn_spatial_bwd=attn_spatial_bwd,
           attn_temporal_bwd=attn_temporal_bwd, adamw_step=adamw_step, gemm_act_dual=gemm_act_dual, gemm_dgelu=gemm_dgelu,
           sepconv_fused=sepconv_fused)


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


# ---- conv2 on pixel pairs: a torch mirror of conv3x3_pair_pack_kernel / conv3x3_pair_launch (csrc/gemm_tcgen05.cu) ----
def conv3x3_pair_taps(w_in):
    """[(ky, jblk, s, pair-row shift)] in the kernel's k-block order for an input of width w_in."""
    taps = []
    for ky in range(3):
        off = ky * w_in
        s = off & 1
        for jblk in range(3 if s else 2):
            taps.append((ky, jblk, s, (off - s) // 2 + jblk))
    return taps


def conv3x3_pair_pack_ref(wt, w_in):
    """wt [64, 3, 3, 32] -> [128, taps * 64]: row = (output pixel p of the pair, co), column = (k-block, input pixel j of
    the pair, ci); W[co, ky, kx, ci] with kx = 2 jblk + j - s - p, zero where kx is outside the filter."""
    taps = conv3x3_pair_taps(w_in)
    out = torch.zeros(128, len(taps) * 64, dtype=wt.dtype)
    for t, (ky, jblk, s, _) in enumerate(taps):
        for p in range(2):
            for j in range(2):
                kx = 2 * jblk + j - s - p
                if 0 <= kx <= 2:
                    out[p * 64:(p + 1) * 64, t * 64 + j * 32:t * 64 + (j + 1) * 32] = wt[:, ky, kx, :]
    return out


def conv3x3_pair_ref(x, wt, bias, act=1):
    """The kernel's data movement in torch: the NHWC input as a matrix of pixel pairs [n h w / 2, 64], one GEMM per k-block
    on the matrix shifted down by the block's pair-row shift (zero fill past the end), outputs on the input's pixel grid
    with the junk columns / rows dropped."""
    n, h, w, cin = x.shape
    assert cin == 32 and (n * h * w) % 2 == 0
    taps = conv3x3_pair_taps(w)
    wp = conv3x3_pair_pack_ref(wt.float(), w)
    a = x.float().reshape(-1, 64)                          # pair rows
    m = a.shape[0]
    acc = torch.zeros(m, 128)
    for t, (_, _, _, shift) in enumerate(taps):
        sh = torch.zeros_like(a)
        if shift < m:
            sh[:m - shift] = a[shift:]
        acc += sh @ wp[:, t * 64:(t + 1) * 64].t()
    y = (acc.reshape(m * 2, 64) + bias.float()).reshape(n, h, w, 64)[:, :h - 2, :w - 2]
    return (F.relu(y) if act == 1 else y).to(x.dtype)


# ---- spatial attention, online softmax with a lazy rescale: a torch mirror of attn_spatial_pp_kernel's arithmetic ----
def attn_online_lazy_ref(q, k, v, scale, step=64, raise_log2=8.0, count=None):
    """q, k, v: [items, tokens, 64] fp32 holding bf16 values.  One row's keys are visited in `step`-key steps; the
    reference m_ref (log2 domain, scores times scale * log2 e) is the first step's maximum and is raised only when a
    later step's maximum exceeds it by more than `raise_log2` — then the accumulated output and denominator are
    multiplied by 2^(old - new).  P is rounded to bf16 before the P.V product (fp32 accumulation), the denominator sums
    the unrounded exponentials: csrc/attn_spatial_pp.cuh.  `count` (a list) receives the number of raises."""
    sl2 = scale * 1.4426950408889634
    n = q.shape[1]
    s = torch.einsum("bid,bjd->bij", q, k)                      # fp32 scores
    m_ref = None
    o = torch.zeros_like(q)
    l = torch.zeros(q.shape[0], n)
    raises = 0
    for j0 in range(0, n, step):
        sj = s[:, :, j0:j0 + step]
        mx = sj.max(dim=-1).values
        if m_ref is None:
            m_ref = mx.clone()
            corr = torch.ones_like(mx)
        else:
            up = (mx - m_ref) * sl2 > raise_log2
            corr = torch.where(up, torch.exp2((m_ref - mx) * sl2), torch.ones_like(mx))
            m_ref = torch.where(up, mx, m_ref)
            raises += int(up.sum())
        p = torch.exp2(sj * sl2 - (m_ref * sl2)[..., None])
        o = o * corr[..., None] + torch.einsum("bij,bjd->bid", p.to(torch.bfloat16).float(), v[:, j0:j0 + step])
        l = l * corr + p.sum(dim=-1)
    if count is not None:
        count.append(raises)
    lse = m_ref * sl2 + torch.log2(l)
    return o / l[..., None], lse
