"""Per-kernel parity checks: each CUDA kernel (through the C ABI) vs a plain PyTorch fp32 reference of the
same op.  Used by tests/test_kernels_gpu.py (pytest, `-m gpu`) and by tools/gpu_check.py, which runs every
check in its own subprocess with a timeout so that one faulting kernel cannot take the others down.

Tolerances (norm-wise: max|a-b| / max|ref|): 1e-5 for fp32 SIMT kernels (5e-5 where exp/erf enter),
1e-2 for bf16-output kernels (one bf16 rounding of the output = 2^-9 relative, plus bf16 inputs).
"""
from __future__ import annotations

import importlib
import math
import os

import torch
import torch.nn.functional as F

from helpers import pkg, rel_err

DEV = "cuda"
TOL_F32 = 2e-5
TOL_BF16 = 1e-2


def _ops():
    return pkg().ops


def _rand(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def _noTF32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _describe_mismatch(got: torch.Tensor, want: torch.Tensor, tol: float) -> str:
    """Structure of the error, to debug descriptor / layout bugs from a single remote run."""
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs()
    thr = tol * want.abs().max().item()
    bad = err > thr
    lines = [f"shape={tuple(got.shape)} bad={int(bad.sum())}/{bad.numel()} maxerr={err.max().item():.4g} "
             f"ref_absmax={want.abs().max().item():.4g} got_absmax={got.abs().max().item():.4g} "
             f"nan={int(torch.isnan(got).sum())}"]
    if got.dim() == 2 and bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        lines.append(f"bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        lines.append(f"bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        r0, c0 = int(rows[0]), int(cols[0])
        lines.append(f"got[{r0},{c0}:{c0 + 8}]={got[r0, c0:c0 + 8].tolist()}")
        lines.append(f"ref[{r0},{c0}:{c0 + 8}]={want[r0, c0:c0 + 8].tolist()}")
        lines.append(f"row-mod-8 bad histogram={[int(bad[i::8].sum()) for i in range(8)]}")
        lines.append(f"col-mod-64/8 bad histogram={[int(bad[:, j::64].sum()) for j in range(0, 64, 8)]}")
    return "\n".join(lines)


def _assert_close(name, got, want, tol):
    r = rel_err(got, want)
    if not (r <= tol):
        raise AssertionError(f"{name}: rel err {r:.3e} > {tol:.1e}\n" + _describe_mismatch(got, want, tol))
    return r


# ------------------------------------------------------------------------------------------------
def check_layernorm():
    ops = _ops()
    out = {}
    for rows, dim in ((37, 728), (1000, 728), (5, 64), (300, 1024), (33, 8)):
        x = _rand(rows, dim, seed=rows) * 3 + 0.5
        g, b = _rand(dim, seed=1) * 0.2 + 1, _rand(dim, seed=2) * 0.1
        ref = F.layer_norm(x, (dim,), g, b, 1e-5)
        out[f"f32_{rows}x{dim}"] = _assert_close("ln f32", ops.layernorm(x, g, b, torch.float32), ref, TOL_F32)
        out[f"bf16_{rows}x{dim}"] = _assert_close("ln bf16", ops.layernorm(x, g, b, torch.bfloat16), ref, TOL_BF16)
        xb = x.to(torch.bfloat16)
        refb = F.layer_norm(xb.float(), (dim,), g, b, 1e-5)
        out[f"bf16in_{rows}x{dim}"] = _assert_close("ln bf16-in", ops.layernorm(xb, g, b, torch.bfloat16), refb, TOL_BF16)
    return out


def check_layernorm_diff():
    ops = _ops()
    out = {}
    for (b, f, p, d) in ((2, 7, 362, 728), (1, 33, 362, 728), (3, 2, 5, 64), (1, 1, 3, 128)):
        x = _rand(b, f, p, d, seed=f) * 2 + 0.3
        g, be = _rand(d, seed=1) * 0.2 + 1, _rand(d, seed=2) * 0.1
        xn_ref = F.layer_norm(x, (d,), g, be, 1e-5)
        diff_ref = torch.cat((xn_ref[:, :2], xn_ref[:, 2:] - xn_ref[:, 1:-1]), dim=1)   # module.py:192
        for dt, tol in ((torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)):
            xn, diff = ops.layernorm_diff(x, g, be, dt)
            out[f"xn_{dt}_{f}"] = _assert_close("ln_diff xn", xn, xn_ref, tol)
            out[f"diff_{dt}_{f}"] = _assert_close("ln_diff diff", diff, diff_ref, tol)
    return out


def _gemm_ref(a, w, bias, residual, act):
    y = a.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.double()
    return y.float()


def _check_gemm_case(m, n, k, bias, residual, act, out_dtype, seed=0):
    ops = _ops()
    a = _rand(m, k, seed=seed, dtype=torch.bfloat16)
    w = _rand(n, k, seed=seed + 1, scale=1 / math.sqrt(k), dtype=torch.bfloat16)
    bi = _rand(n, seed=seed + 2) if bias else None
    res = _rand(m, n, seed=seed + 3) if residual else None
    ref = _gemm_ref(a, w, bi, res, act)
    if residual and out_dtype == torch.float32:
        out = res.clone()
        got = ops.gemm(a, w, bias=bi, residual=out, act=act, out=out)     # in-place residual update
    else:
        got = ops.gemm(a, w, bias=bi, residual=res, act=act, out_dtype=out_dtype)
    torch.cuda.synchronize()
    tol = 2e-3 if out_dtype == torch.float32 else TOL_BF16
    return _assert_close(f"gemm m={m} n={n} k={k} bias={bias} res={residual} act={act} out={out_dtype}",
                         got.view(m, n), ref, tol)


def check_gemm_basic():
    """Smallest tcgen05 cases first: one tile, one k-block; then multi-k, multi-tile."""
    out = {}
    out["128x256x64"] = _check_gemm_case(128, 256, 64, False, False, 0, torch.float32)
    out["128x256x256"] = _check_gemm_case(128, 256, 256, False, False, 0, torch.float32)
    out["256x512x128"] = _check_gemm_case(256, 512, 128, False, False, 0, torch.bfloat16)
    out["128x64x64"] = _check_gemm_case(128, 64, 64, False, False, 0, torch.float32)
    out["128x128x64"] = _check_gemm_case(128, 128, 64, False, False, 0, torch.float32)
    return out


def check_gemm_shapes():
    """Every (N, K) of the path with ragged M, all epilogues."""
    out = {}
    m = 2534 + 77
    out["to_qk"] = _check_gemm_case(m, 1024, 728, False, False, 0, torch.bfloat16, 1)
    out["to_v"] = _check_gemm_case(m, 512, 728, False, False, 0, torch.bfloat16, 2)
    out["to_out"] = _check_gemm_case(m, 728, 512, True, False, 0, torch.bfloat16, 3)
    out["to_qkv"] = _check_gemm_case(m, 1536, 728, False, False, 0, torch.bfloat16, 4)
    out["s_out"] = _check_gemm_case(m, 728, 512, True, True, 0, torch.float32, 5)
    out["ff1"] = _check_gemm_case(m, 2912, 728, True, False, 2, torch.bfloat16, 6)
    out["ff2"] = _check_gemm_case(m, 728, 2912, True, True, 0, torch.float32, 7)
    # entry-flow pointwise / skip shapes (bias = folded BN shift, optional ReLU)
    for (n, k) in ((128, 64), (128, 128), (256, 128), (256, 256), (728, 256), (728, 728)):
        out[f"pw_{k}_{n}"] = _check_gemm_case(1000 + n, n, k, True, False, 1, torch.bfloat16, n + k)
    out["k32"] = _check_gemm_case(300, 64, 32, True, False, 1, torch.bfloat16, 9)
    out["many_tiles"] = _check_gemm_case(128 * 200 + 5, 728, 728, True, False, 0, torch.bfloat16, 10)
    # >= 6 tiles per cluster with a ragged last N tile whose upper column groups are empty (N = 11*256 + 96, like
    # ff1): exercises the TMEM double-buffer hand-off on every path (a missed tmem_empty arrival deadlocks here)
    out["ragged_n_many_tiles"] = _check_gemm_case(256 * 40 + 3, 2912, 128, True, False, 2, torch.bfloat16, 11)
    out["ragged_n_many_tiles_f32"] = _check_gemm_case(256 * 120 + 3, 1120, 64, True, True, 0, torch.float32, 12)
    return out


def check_gemm_lnfold():
    """LayerNorm folded around a GEMM pair (istvt_gemm_rowstats_fwd + istvt_gemm_lnfold_fwd, engine.fold_layernorm)
    vs the unfused definition Linear(LayerNorm(y)) evaluated in fp64 on the same bf16-rounded y; ragged M, the path's
    shapes (to_out 512 -> 728, to_qkv 728 -> 1536), a row-mean several sigma away from zero, near-constant rows."""
    ops = _ops()
    eng = importlib.import_module("2023-tifs-istvt_b200.engine")
    out = {}
    for m, shift in ((1000, 0.0), (257, 3.0), (4099, -1.5)):
        a = _rand(m, 512, seed=m).to(torch.bfloat16)
        w1 = (_rand(728, 512, seed=m + 1) * 512 ** -0.5).to(torch.bfloat16)
        b1 = _rand(728, seed=m + 2) * 0.5 + shift
        gamma, beta = _rand(728, seed=m + 3).abs() + 0.3, _rand(728, seed=m + 4) * 0.4
        w2 = _rand(1536, 728, seed=m + 5) * 728 ** -0.5
        a[-3:] *= 1e-3                      # rows whose variance is tiny next to their mean (= the bias)
        stats = torch.full((m, 12, 2), float("nan"), device="cuda")
        y = ops.gemm_rowstats(a, w1, b1, stats)
        yref = _gemm_ref(a, w1, b1, None, 0)
        out[f"y_{m}"] = _assert_close("lnfold y", y, yref, TOL_BF16)
        yd = yref.double()                 # the statistics are taken before the bf16 rounding
        blocks = [yd[:, 64 * g: 64 * g + 64] for g in range(12)]
        want_stats = torch.stack([torch.stack((b.sum(1), ((b - b.mean(1, keepdim=True)) ** 2).sum(1)), 1) for b in blocks], 1)
        out[f"stats_{m}"] = _assert_close("lnfold stats", stats, want_stats.float(), 2e-5)
        mr = ops.ln_stats_finalize(stats, 728)
        want_mr = torch.stack((yd.mean(1), torch.rsqrt(yd.var(1, unbiased=False) + 1e-5)), 1)
        out[f"mu_rstd_{m}"] = _assert_close("lnfold mu/rstd", mr, want_mr.float(), 2e-5)
        y2 = ops.gemm_rowstats(a, w1, b1, torch.empty_like(stats))
        assert torch.equal(y, y2), "the statistics output must not change the GEMM result"
        wf, c, d = eng.fold_layernorm(w2, gamma, beta, torch.bfloat16)
        z = ops.gemm_lnfold(y, wf, mr, c, d)
        zref = torch.nn.functional.layer_norm(y.double(), (728,), gamma.double(), beta.double(), 1e-5) @ w2.double().t()
        out[f"z_{m}"] = _assert_close("lnfold z", z, zref.float(), TOL_BF16)
    return out


def check_gemm_mlp_fusions():
    """Training-step MLP fusions vs fp64: istvt_gemm_act_dual_fwd (pre-activation and GELU from one accumulator) and
    istvt_gemm_dgelu_fwd ((A W^T) o gelu'(pre)); ragged M and N, row-pitched A / W operands (K = 728 at pitch 768)."""
    ops = _ops()
    out = {}
    for m in (1000, 257, 4099):
        a = ops.empty_rows((m, 728), torch.bfloat16, "cuda")
        a.copy_(_rand(m, 728, seed=m))
        w1 = ops.pad_rows((_rand(2912, 728, seed=m + 1) * 728 ** -0.5).to(torch.bfloat16))
        b1 = _rand(2912, seed=m + 2) * 0.5
        hid, pre = ops.gemm_act_dual(a, w1, b1, ops.ACT_GELU)
        pre_ref = a.double() @ w1.double().t() + b1.double()
        out[f"pre_{m}"] = _assert_close("dual pre", pre, pre_ref.float(), TOL_BF16)
        out[f"hid_{m}"] = _assert_close("dual gelu", hid, torch.nn.functional.gelu(pre_ref).float(), TOL_BF16)
        plain = ops.gemm(a, w1, bias=b1, act=ops.ACT_GELU)
        assert torch.equal(plain, hid), "dual-output GELU differs from the single-output epilogue"
        # data gradient through Linear(2912 -> 728)^T and the GELU: g [m, 728] . W2 [728, 2912] o gelu'(pre)
        g = ops.empty_rows((m, 728), torch.bfloat16, "cuda")
        g.copy_(_rand(m, 728, seed=m + 3))
        w2t = ops.pad_rows((_rand(2912, 728, seed=m + 4) * 728 ** -0.5).to(torch.bfloat16))     # = W2^T, [2912, 728]
        got = ops.gemm_dgelu(g, w2t, pre)
        x = pre.double()
        dg = 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327
        ref = (g.double() @ w2t.double().t()) * dg
        out[f"dgelu_{m}"] = _assert_close("dgelu", got, ref.float(), TOL_BF16)
    return out


def check_gemm_f32():
    ops = _ops()
    out = {}
    for (m, n, k, bias, res, act) in ((130, 64, 32, True, False, 1), (1000, 728, 512, True, True, 0),
                                      (777, 2912, 728, True, False, 2), (513, 1024, 728, False, False, 0)):
        a = _rand(m, k, seed=m)
        w = _rand(n, k, seed=n, scale=1 / math.sqrt(k))
        bi = _rand(n, seed=3) if bias else None
        r = _rand(m, n, seed=4) if res else None
        ref = _gemm_ref(a, w, bi, r, act)
        got = ops.gemm(a, w, bias=bi, residual=r, act=act)
        out[f"{m}x{n}x{k}"] = _assert_close("gemm_f32", got, ref, TOL_F32)
    return out


def check_conv3x3():
    ops = _ops()
    _noTF32()
    out = {}
    for dt, tol in ((torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)):
        for (n, h, w, cin, cout) in ((2, 9, 11, 32, 64), (1, 149, 149, 32, 64), (3, 20, 17, 32, 64)):
            x = _rand(n, h, w, cin, seed=h).to(dt)
            wt = (_rand(cout, 3, 3, cin, seed=5) / math.sqrt(9 * cin)).to(dt)
            b = _rand(cout, seed=6) * 0.1
            ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), b)).permute(0, 2, 3, 1)
            got = ops.conv3x3(x, wt, b, act=1)
            out[f"{dt}_{n}x{h}x{w}"] = _assert_close(f"conv3x3 {dt} {n}x{h}x{w}", got.reshape(-1, cout),
                                                     ref.reshape(-1, cout), tol)
    # without the ReLU (training mode: BatchNorm batch statistics follow), and the data-gradient shape 64 -> 32 channels
    for (cin, cout) in ((32, 64), (64, 32)):
        x = _rand(2, 13, 12, cin, seed=cin).to(torch.bfloat16)
        wt = (_rand(cout, 3, 3, cin, seed=7) / math.sqrt(9 * cin)).to(torch.bfloat16)
        b = _rand(cout, seed=8) * 0.1
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), b).permute(0, 2, 3, 1)
        got = ops.conv3x3(x, wt, b, act=0)
        out[f"noact_{cin}_{cout}"] = _assert_close(f"conv3x3 no act {cin}->{cout}", got.reshape(-1, cout),
                                                   ref.reshape(-1, cout), TOL_BF16)
    # conv2's shape (32 -> 64) has four kernels: the pixel-pair formulation (default when the number of input pixels is
    # even; odd and even input widths take 7 / 6 k-blocks), the 9-tap TMA formulation (every other shape), the strip
    # kernel (conv3x3_strip.cu) and the gathered-operand kernel (conv3x3_tc.cu) — the last two measured slower and kept
    # as A/B alternatives: same answers from all of them, ragged strips / row blocks / last tiles included, several items
    # per CTA, 1 x 1 outputs
    prev = os.environ.get("ISTVT_CONV2_KERNEL")
    try:
        for kern in ("pair", "strip", "taps", "gather"):
            os.environ["ISTVT_CONV2_KERNEL"] = kern
            for (n, h, w, act) in ((2, 13, 12, 0), (1, 149, 149, 1), (2, 149, 149, 1), (40, 41, 78, 1), (3, 3, 3, 0),
                                   (2, 3, 3, 1), (4, 5, 4, 0), (6, 37, 51, 1)):
                x = _rand(n, h, w, 32, seed=h).to(torch.bfloat16)
                wt = (_rand(64, 3, 3, 32, seed=7) / math.sqrt(288)).to(torch.bfloat16)
                b = _rand(64, seed=8) * 0.1
                ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), b)
                ref = (ref.relu() if act else ref).permute(0, 2, 3, 1)
                got = ops.conv3x3(x, wt, b, act=act, kernel=kern)
                out[f"{kern}_{n}x{h}x{w}"] = _assert_close(f"conv3x3 {kern} {n}x{h}x{w}", got.reshape(-1, 64),
                                                           ref.reshape(-1, 64), TOL_BF16)
    finally:
        if prev is None:
            os.environ.pop("ISTVT_CONV2_KERNEL", None)
        else:
            os.environ["ISTVT_CONV2_KERNEL"] = prev
    # the CUDA weight rearrangement of the pixel-pair kernel against its torch mirror (tests/torch_ops.py), bit for bit
    import torch_ops
    for w_in in (149, 78):
        wt = (_rand(64, 3, 3, 32, seed=w_in) / math.sqrt(288)).to(torch.bfloat16)
        wp = ops.conv3x3_pair_weights(wt, w_in)
        torch.cuda.synchronize()
        assert torch.equal(wp.cpu(), torch_ops.conv3x3_pair_pack_ref(wt.cpu(), w_in)), f"conv3x3_pair_pack w_in={w_in}"
        assert ops.conv3x3_pair_weights(wt, w_in) is wp, "pair weights must be memoised on the weight tensor"
    return out


def check_conv_stem():
    ops = _ops()
    _noTF32()
    out = {}
    for (n, h, w) in ((2, 300, 300), (1, 299, 299), (3, 31, 40)):
        x = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(h)).to(DEV) * 2 - 1
        wt = _rand(32, 3, 3, 3, seed=1) * 0.3
        b = _rand(32, seed=2) * 0.1
        ref = F.relu(F.conv2d(x, wt, b, stride=2)).permute(0, 2, 3, 1).contiguous()
        out[f"f32_{h}"] = _assert_close("stem f32", ops.conv_stem(x, wt, b, torch.float32), ref, TOL_F32)
        out[f"bf16_{h}"] = _assert_close("stem bf16", ops.conv_stem(x, wt, b, torch.bfloat16), ref, TOL_BF16)
    return out


def check_conv_stem_u8():
    """uint8 NHWC stem (normalisation folded into the weights by the host) vs torch conv on the normalised clip."""
    ops = _ops()
    _noTF32()
    import importlib
    eng = importlib.import_module("2023-tifs-istvt_b200.engine")
    out = {}
    for (n, h, w) in ((2, 300, 300), (1, 299, 299), (3, 7, 9), (1, 8, 6), (2, 5, 5)):   # even / odd row pitch, ragged pairs
        g = torch.Generator().manual_seed(h * 31 + w)
        u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8).to(DEV)
        wt = _rand(32, 3, 3, 3, seed=1) * 0.2
        b = _rand(32, seed=2) * 0.1
        mean, std = (0.5, 0.4, 0.45), (0.5, 0.25, 0.3)
        xn = (u8.float() / 255.0 - torch.tensor(mean, device=DEV)) / torch.tensor(std, device=DEV)
        ref = F.relu(F.conv2d(xn.permute(0, 3, 1, 2), wt, b, stride=2)).permute(0, 2, 3, 1).contiguous()
        w2, b2 = eng.fold_input_norm(wt, b, mean, std)
        out[f"f32_{h}x{w}"] = _assert_close("stem u8 f32", ops.conv_stem_u8(u8, w2, b2, torch.float32), ref, 2e-5)
        out[f"bf16_{h}x{w}"] = _assert_close("stem u8 bf16", ops.conv_stem_u8(u8, w2, b2, torch.bfloat16), ref, TOL_BF16)
    return out


def check_xception_tail():
    ops = _ops()
    _noTF32()
    out = {}
    for dt, tol in ((torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)):
        a = _rand(3, 19, 19, 728, seed=1).to(dt)
        b = _rand(3, 19, 19, 728, seed=2).to(dt)
        out[f"add_{dt}"] = _assert_close("add", ops.add(a, b), a.float() + b.float(), tol)
        x = _rand(5, 10, 10, 2048, seed=3).to(dt)
        w = _rand(2, 2048, seed=4) * 0.05
        bias = _rand(2, seed=5)
        ref = F.linear(F.relu(x.float()).mean(dim=(1, 2)), w, bias)
        out[f"pool_linear_{dt}"] = _assert_close("pool_linear", ops.pool_linear(x, w, bias), ref, 1e-5 if dt == torch.float32 else 2e-3)
        x = _rand(2, 3, 7, 24, seed=6).to(dt)
        w = _rand(5, 24, seed=7)
        ref = F.linear(x.float().mean(dim=(1, 2)), w, None)
        out[f"pool_linear_norelu_{dt}"] = _assert_close("pool_linear", ops.pool_linear(x, w, None, relu=False), ref, 1e-5 if dt == torch.float32 else 2e-3)
    return out


def check_dwconv():
    ops = _ops()
    _noTF32()
    out = {}
    for (n, h, w, c, relu) in ((2, 147, 147, 64, False), (1, 74, 74, 128, True), (2, 37, 37, 728, True),
                               (1, 5, 3, 8, False), (1, 1, 1, 16, True), (1, 9, 2, 256, False),
                               (20, 37, 37, 728, True), (3, 40, 77, 72, False)):   # several items per CTA; ragged strips
        x = _rand(n, h, w, c, seed=c)
        wt = _rand(3, 3, c, seed=7) * 0.3
        xin = F.relu(x) if relu else x
        ref = F.conv2d(xin.permute(0, 3, 1, 2), wt.permute(2, 0, 1).unsqueeze(1), None, 1, 1, 1, groups=c)
        ref = ref.permute(0, 2, 3, 1).contiguous()
        out[f"f32_{h}_{c}"] = _assert_close("dw f32", ops.dwconv3x3(x, wt, relu), ref, TOL_F32)
        xb = x.to(torch.bfloat16)
        xinb = F.relu(xb.float()) if relu else xb.float()
        refb = F.conv2d(xinb.permute(0, 3, 1, 2), wt.permute(2, 0, 1).unsqueeze(1), None, 1, 1, 1, groups=c)
        refb = refb.permute(0, 2, 3, 1).contiguous()
        out[f"bf16_{h}_{c}"] = _assert_close("dw bf16", ops.dwconv3x3(xb, wt, relu), refb, TOL_BF16)
    return out


def check_sepconv_fused():
    """istvt_sepconv_fused_fwd (depthwise 3x3 + pointwise 1x1 + bias + ReLU, depthwise result kept on chip) vs the two-kernel
    path it replaces (istvt_dwconv3x3_fwd + istvt_gemm_fwd: same bf16 rounding of the depthwise result, so the two must
    agree to the last bits of the fp32 accumulation order) and vs an fp32 torch reference; the entry flow's shapes, ragged
    strips / row blocks, several items per CTA, image borders, all four (relu_in, act) combinations."""
    ops = _ops()
    _noTF32()
    out = {}
    cases = ((2, 147, 147, 64, 128, False, 1), (1, 147, 147, 128, 128, False, 0), (2, 74, 74, 128, 256, True, 1),
             (1, 13, 12, 64, 64, True, 0), (3, 20, 17, 192, 192, False, 1), (1, 7, 40, 256, 128, True, 1),
             (5, 37, 37, 128, 256, True, 0), (1, 1, 1, 64, 128, False, 1), (40, 74, 74, 64, 128, True, 1))
    for (n, h, w, c, n_out, relu_in, act) in cases:
        assert ops.sepconv_fused_supported(c, n_out, w)
        x = _rand(n, h, w, c, seed=c + h).to(torch.bfloat16)
        dw = _rand(3, 3, c, seed=7) * 0.3
        pw = (_rand(n_out, c, seed=8) * c ** -0.5).to(torch.bfloat16)
        bias = _rand(n_out, seed=9) * 0.2
        got = ops.sepconv_fused(x, dw, pw, bias, relu_in, act)
        two = ops.gemm(ops.dwconv3x3(x, dw, relu_in), pw, bias=bias, act=act).view(n, h, w, n_out)
        xin = F.relu(x.float()) if relu_in else x.float()
        d = F.conv2d(xin.permute(0, 3, 1, 2), dw.permute(2, 0, 1).unsqueeze(1), None, 1, 1, 1, groups=c)
        d = d.permute(0, 2, 3, 1).to(torch.bfloat16).float()             # the depthwise result is a bf16 operand
        ref = d.reshape(-1, c) @ pw.float().t() + bias
        ref = (F.relu(ref) if act else ref).view(n, h, w, n_out)
        tag = f"{n}x{h}x{w}_{c}_{n_out}"
        out["ref_" + tag] = _assert_close("sepconv_fused vs torch " + tag, got, ref, TOL_BF16)
        out["two_" + tag] = _assert_close("sepconv_fused vs dwconv + gemm " + tag, got, two.float(), 4e-3)
    assert not ops.sepconv_fused_supported(256, 256, 74) and not ops.sepconv_fused_supported(728, 728, 37)
    return out


def check_pool_subsample_tokens():
    ops = _ops()
    out = {}
    for dt, tol in ((torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)):
        for (n, h, w, c) in ((2, 147, 147, 128), (3, 74, 74, 256), (2, 37, 37, 728), (1, 4, 7, 8)):
            x = _rand(n, h, w, c, seed=h).to(dt)
            sub = ops.subsample2(x)
            assert torch.equal(sub, x[:, ::2, ::2].contiguous()), "subsample2 must be an exact gather"
            skip = _rand(*sub.shape, seed=3).to(dt)
            ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1) + skip.float()
            out[f"pool_{dt}_{h}"] = _assert_close("pool_add", ops.pool_add(x, skip), ref, tol)
        # block-3 variant writing tokens
        b, t, h, w, c = 2, 3, 37, 37, 728
        x = _rand(b * t, h, w, c, seed=11).to(dt)
        skip = _rand(b * t, 19, 19, c, seed=12).to(dt)
        pos = _rand(t, 362, c, seed=13)
        space, temporal = _rand(c, seed=14), _rand(c, seed=15)
        tokens = torch.full((b, t + 1, 362, c), float("nan"), device=DEV)
        ops.pool_add_tokens(x, skip, pos, tokens, b, t)
        ops.token_fill(tokens, space, temporal, pos)
        pooled = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1) + skip.float()
        ref = torch.empty_like(tokens)
        ref[:, 0] = temporal
        ref[:, 1:, 0] = space + pos[:, 0]
        ref[:, 1:, 1:] = pooled.reshape(b, t, 361, c) + pos[:, 1:]
        out[f"tokens_{dt}"] = _assert_close("tokens", tokens, ref, tol)
    return out


def _attn_ref(q, k, v, scale):
    a = torch.softmax(torch.matmul(q.double(), k.double().transpose(-1, -2)) * scale, dim=-1)
    return torch.matmul(a, v.double()).float(), a.float()


def check_attn_temporal():
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    # bf16: f <= 8 runs the one-tile warp-per-(clip, position, head) mma.sync kernel, 9 <= f <= 48 the frame-tiled one
    # (1-3 m16 tiles; f = 17 / 33 = 16 MT + 1 — T = 16 / 32 frames + the temporal class frame, the long-clip
    # configuration — the variant that does the last frame on the FMA pipe); fp32 (validation mode) the SIMT kernel, f <= 36
    for (b, f, p) in ((2, 7, 362), (1, 33, 362), (1, 2, 5), (3, 8, 101), (1, 1, 9), (2, 9, 37), (1, 16, 50), (2, 17, 19),
                      (2, 32, 21), (3, 33, 5), (1, 34, 7), (1, 47, 11), (1, 48, 13)):
        for dt, tol in ((torch.float32, 5e-5), (torch.bfloat16, TOL_BF16)):
            if dt == torch.float32 and f > 36:
                continue
            rows = b * f * p
            qk = (_rand(rows, 1024, seed=f) * 1.5).to(dt)
            v = _rand(rows, 512, seed=f + 1).to(dt)
            o, probs = ops.attn_temporal(qk, v, b, f, p, heads, scale, want_probs=True)
            split = lambda t: t.float().reshape(b, f, p, heads, 64).permute(0, 3, 2, 1, 4)   # b h p f d
            q, k = split(qk[:, :512]), split(qk[:, 512:])
            oref, aref = _attn_ref(q, k, split(v), scale)
            oref = oref.permute(0, 3, 2, 1, 4).reshape(rows, 512)
            out[f"out_{dt}_{f}"] = _assert_close("attn_t out", o, oref, tol)
            out[f"probs_{dt}_{f}"] = _assert_close("attn_t probs", probs, aref, 5e-5 if dt == torch.float32 else 2e-3)
            o2, none = ops.attn_temporal(qk, v, b, f, p, heads, scale, want_probs=False)
            assert none is None and torch.equal(o, o2), "probs emission must not change the output"
    try:      # beyond the kernels' range: a loud error, not a wrong answer
        ops.attn_temporal(torch.zeros(49 * 8, 1024, device="cuda", dtype=torch.bfloat16),
                          torch.zeros(49 * 8, 512, device="cuda", dtype=torch.bfloat16), 1, 49, 8, heads, scale)
        raise AssertionError("49 frames must be rejected")
    except RuntimeError as e:
        assert "not supported" in str(e), e
    return out


def check_attn_spatial_f32():
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    for (bf, p) in ((3, 362), (2, 50)):
        qkv = _rand(bf * p, 1536, seed=p) * 1.5
        o, probs = ops.attn_spatial(qkv, bf, p, heads, scale, want_probs=True)
        split = lambda t: t.reshape(bf, p, heads, 64).permute(0, 2, 1, 3)
        oref, aref = _attn_ref(split(qkv[:, :512]), split(qkv[:, 512:1024]), split(qkv[:, 1024:]), scale)
        out[f"out_{p}"] = _assert_close("attn_s f32 out", o, oref.permute(0, 2, 1, 3).reshape(bf * p, 512), 5e-5)
        out[f"probs_{p}"] = _assert_close("attn_s f32 probs", probs, aref, 5e-5)
    return out


def check_attn_spatial_bf16():
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    # (40, 362): 320 (frame, head) items on 148 persistent CTAs -> the cross-item prefetch / ring wrap-around paths
    # amp 3.0: logits with a std of 13 in the exp2 domain (near one-hot rows)
    for (bf, p, amp) in ((1, 362, 1.0), (5, 362, 2.0), (2, 128, 1.0), (2, 200, 1.0), (3, 50, 1.0), (40, 362, 1.5),
                         (21, 130, 1.0), (45, 384, 1.0), (4, 362, 3.0), (3, 257, 3.0)):
        qkv = (_rand(bf * p, 1536, seed=p + bf) * amp).to(torch.bfloat16)
        o, probs = ops.attn_spatial(qkv, bf, p, heads, scale, want_probs=True)
        torch.cuda.synchronize()
        split = lambda t: t.float().reshape(bf, p, heads, 64).permute(0, 2, 1, 3)
        oref, aref = _attn_ref(split(qkv[:, :512]), split(qkv[:, 512:1024]), split(qkv[:, 1024:]), scale)
        name = f"{bf}x{p}"
        out[f"probs_{name}"] = _assert_close(f"attn_s bf16 probs {name}", probs.reshape(-1, p), aref.reshape(-1, p), 2e-3)
        out[f"out_{name}"] = _assert_close(f"attn_s bf16 out {name}", o,
                                           oref.permute(0, 2, 1, 3).reshape(bf * p, 512), 1.5e-2)
        # production kernel (persistent, P in TMEM) vs the attention-map kernel: same math, different schedule
        o2, none = ops.attn_spatial(qkv, bf, p, heads, scale, want_probs=False)
        torch.cuda.synchronize()
        assert none is None
        out[f"pipe_{name}"] = _assert_close(f"attn_s bf16 pipelined out {name}", o2,
                                            oref.permute(0, 2, 1, 3).reshape(bf * p, 512), 1.5e-2)
        out[f"pipe_vs_map_{name}"] = _assert_close(f"attn_s bf16 pipelined vs map kernel {name}", o2, o, 8e-3)
        # both persistent kernels forced (the default picks by mode: online-softmax kernel for inference, exact-max kernel
        # when the log-sum-exp is wanted), each also in its lse mode
        lse_ref = torch.logsumexp(torch.matmul(split(qkv[:, :512]), split(qkv[:, 512:1024]).transpose(-1, -2)) * scale,
                                  dim=-1) / math.log(2.0)
        prev = os.environ.get("ISTVT_SA_KERNEL")
        try:
            for kern in ("pp3", "pp", "pipe"):       # pp3 has no lse mode (the launcher takes pp for it)
                os.environ["ISTVT_SA_KERNEL"] = kern
                o3, _ = ops.attn_spatial(qkv, bf, p, heads, scale, want_probs=False)
                o4, lse = ops.attn_spatial_lse(qkv, bf, p, heads, scale)
                torch.cuda.synchronize()
                out[f"{kern}_{name}"] = _assert_close(f"attn_s bf16 {kern} kernel out {name}", o3,
                                                      oref.permute(0, 2, 1, 3).reshape(bf * p, 512), 1.5e-2)
                if kern != "pp3":
                    assert torch.equal(o3, o4), f"{kern} kernel: lse mode changes the output ({name})"
                out[f"{kern}_lse_{name}"] = _assert_close(f"attn_s bf16 {kern} kernel lse {name}", lse, lse_ref.float(), 2e-3)
        finally:
            if prev is None:
                os.environ.pop("ISTVT_SA_KERNEL", None)
            else:
                os.environ["ISTVT_SA_KERNEL"] = prev
    return out


def check_attn_spatial_spiky():
    """Rows whose softmax is dominated by ONE key by a wide margin (~30, ~60 and ~200 nats: near one-hot to exactly
    one-hot in fp32): the two-pass kernel subtracts the exact row maximum, so these must match the reference and the
    log-sum-exp handed to the backward must stay exact.  (A variant that took the stabiliser from a sample of the keys
    was measured and dropped, profiles/README.md r3y; this check is what it had to pass.)"""
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    for (bf, p, key, boost) in ((3, 362, 13, 30.0), (3, 362, 300, 30.0), (2, 362, 13, 200.0), (2, 200, 77, 200.0),
                                (2, 362, 361, 60.0)):
        qkv = _rand(bf * p, 1536, seed=key + p).reshape(bf, p, 3, heads, 64)
        # every query gets the component 8 along dim 0; the spiked key gets boost / (8 * scale) there:
        # logit(query, key) = 8 * boost / (8 * scale) * scale + noise = boost nats above the rest
        qkv[:, :, 0, :, 0] = 8.0
        qkv[:, :, 1, :, 0] = 0.0
        qkv[:, key, 1, :, 0] = boost / (8.0 * scale)
        qkv = qkv.reshape(bf * p, 1536).to(torch.bfloat16)
        split = lambda t: t.float().reshape(bf, p, heads, 64).permute(0, 2, 1, 3)
        oref, _ = _attn_ref(split(qkv[:, :512]), split(qkv[:, 512:1024]), split(qkv[:, 1024:]), scale)
        s = torch.einsum("bhid,bhjd->bhij", split(qkv[:, :512]), split(qkv[:, 512:1024])) * scale
        lse_ref = torch.logsumexp(s, dim=-1) * 1.4426950408889634       # log2 domain
        # both persistent kernels: the exact-max one, and the online-softmax one, whose reference is raised in the middle
        # of a row here (spikes in the first / second half of a 128-key chunk, in the first and in later chunks: the
        # rescale of O, of the half-written P chunk and of the denominator)
        prev = os.environ.get("ISTVT_SA_KERNEL")
        try:
            for kern in ("pp3", "pp", "pipe"):
                os.environ["ISTVT_SA_KERNEL"] = kern
                o, none = ops.attn_spatial(qkv, bf, p, heads, scale, want_probs=False)
                torch.cuda.synchronize()
                assert torch.isfinite(o.float()).all(), "non-finite attention output"
                out[f"spike_{kern}_{p}_{key}_{int(boost)}"] = _assert_close(
                    f"attn_s spiky {kern} key={key} boost={boost}", o, oref.permute(0, 2, 1, 3).reshape(bf * p, 512), 1.5e-2)
                o2, lse = ops.attn_spatial_lse(qkv, bf, p, heads, scale)          # training forward: same kernel + LSE
                out[f"lse_{kern}_{p}_{key}_{int(boost)}"] = _assert_close("attn_s spiky lse", lse, lse_ref, 2e-3)
        finally:
            if prev is None:
                os.environ.pop("ISTVT_SA_KERNEL", None)
            else:
                os.environ["ISTVT_SA_KERNEL"] = prev
    return out


def check_attn_joint():
    """Key-streaming joint attention (istvt_attn_joint_fwd, module.py:53-63) vs an fp64 softmax(QK^T)V: sequences of
    1 .. 17 key blocks, ragged last blocks, a single-token sequence, near-one-hot logits, and a dominant key that sits
    in a LATE block (every earlier block's partial result must be rescaled away)."""
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    split = lambda t, b, n: t.float().reshape(b, n, heads, 64).permute(0, 2, 1, 3)
    def ref(qkv, b, n):
        o, _ = _attn_ref(split(qkv[:, :512], b, n), split(qkv[:, 512:1024], b, n), split(qkv[:, 1024:], b, n), scale)
        return o.permute(0, 2, 1, 3).reshape(b * n, 512)
    kept = {}
    for (b, n, amp) in ((1, 2167, 1.0), (2, 300, 1.0), (3, 128, 1.0), (2, 129, 2.0), (2, 7, 1.0), (1, 1, 1.0),
                        (2, 1000, 3.0), (1, 256, 1.0), (20, 362, 1.5), (1, 4096, 1.0), (1, 32 * 361 + 1, 1.0)):
        qkv = (_rand(b * n, 1536, seed=n + b) * amp).to(torch.bfloat16)
        o = ops.attn_joint(qkv, b, n, heads, scale)
        torch.cuda.synchronize()
        assert torch.isfinite(o.float()).all(), f"non-finite joint attention output at {b}x{n}"
        want = ref(qkv, b, n)
        out[f"bf16_{b}x{n}"] = _assert_close(f"attn_joint bf16 {b}x{n}", o, want, 1.5e-2)
        kept[(b, n)] = (qkv, o)
    # a dominant key in block 11 of 17 (and one in block 0): online rescaling across blocks
    for (n, key, boost) in ((2167, 1500, 40.0), (2167, 5, 40.0), (700, 699, 200.0)):
        qkv = _rand(2 * n, 1536, seed=key).reshape(2, n, 3, heads, 64)
        qkv[:, :, 0, :, 0] = 8.0
        qkv[:, :, 1, :, 0] = 0.0
        qkv[:, key, 1, :, 0] = boost / (8.0 * scale)
        qkv = qkv.reshape(2 * n, 1536).to(torch.bfloat16)
        o = ops.attn_joint(qkv, 2, n, heads, scale)
        torch.cuda.synchronize()
        assert torch.isfinite(o.float()).all(), "non-finite joint attention output (spiky)"
        out[f"spike_{n}_{key}"] = _assert_close(f"attn_joint spiky n={n} key={key}", o, ref(qkv, 2, n), 1.5e-2)
    # fp32 validation kernel
    for (b, n) in ((1, 2167), (2, 300), (2, 7), (1, 65)):
        qkv = _rand(b * n, 1536, seed=n) * 1.5
        o = ops.attn_joint(qkv, b, n, heads, scale)
        out[f"f32_{b}x{n}"] = _assert_close(f"attn_joint f32 {b}x{n}", o, ref(qkv, b, n), 5e-5)
    # same math as the all-keys-resident spatial kernel wherever that one applies
    for (b, n), (qkv, o) in kept.items():
        if 7 <= n <= 384:
            o2, _ = ops.attn_spatial(qkv, b, n, heads, scale)
            out[f"vs_spatial_{b}x{n}"] = _assert_close(f"attn_joint vs attn_spatial {b}x{n}", o, o2, 8e-3)
    # short sequences through the spatial kernels (ViViT's temporal transformer: 7 tokens per clip)
    for dt, tol in ((torch.float32, 5e-5), (torch.bfloat16, 1.5e-2)):
        qkv = (_rand(5 * 7, 1536, seed=77) * 1.5).to(dt)
        o, _ = ops.attn_spatial(qkv, 5, 7, heads, scale)
        out[f"spatial_7tok_{dt}"] = _assert_close(f"attn_spatial 7 tokens {dt}", o, ref(qkv, 5, 7), tol)
    return out


def check_token_build():
    """Class token + patches (+ positional embedding) assembly of the ablation transformers vs torch cat / add."""
    ops = _ops()
    out = {}
    dim = 728
    for (seqs, n, period, with_pos, dt) in ((12, 361, 6, True, torch.bfloat16), (4, 361, 2, True, torch.float32),
                                            (2, 2166, 1, True, torch.float32), (3, 6, 1, False, torch.float32),
                                            (2, 2166, 1, True, torch.bfloat16)):
        src = _rand(seqs * n, dim, seed=n).to(dt)
        cls = _rand(dim, seed=1)
        pos = _rand(period, n + 1, dim, seed=2) if with_pos else None
        tok = ops.token_build(src, cls, pos, seqs, n, pos_period=period)
        ref = torch.cat((cls.reshape(1, 1, dim).expand(seqs, 1, dim), src.float().reshape(seqs, n, dim)), dim=1)
        if with_pos:
            ref = ref + pos[torch.arange(seqs, device=DEV) % period]
        assert tok.dtype == torch.float32 and tuple(tok.shape) == (seqs, n + 1, dim)
        out[f"{seqs}x{n}_{dt}"] = _assert_close(f"token_build {seqs}x{n}", tok.reshape(-1, dim), ref.reshape(-1, dim), 1e-6)
    return out


def check_head():
    ops = _ops()
    b, f, p, d = 3, 7, 362, 728
    tokens = _rand(b, f, p, d, seed=1) * 2
    ng, nb, hg, hb = _rand(d, seed=2) * 0.2 + 1, _rand(d, seed=3) * 0.1, _rand(d, seed=4) * 0.2 + 1, _rand(d, seed=5) * 0.1
    hw, hbias = _rand(d, seed=6) * 0.05, _rand(1, seed=7)
    x = F.layer_norm(tokens[:, 0, 0], (d,), ng, nb, 1e-5)
    ref = F.layer_norm(x, (d,), hg, hb, 1e-5) @ hw[:, None] + hbias
    got = ops.head(tokens, ng, nb, hg, hb, hw, hbias)
    return {"head": _assert_close("head", got, ref, TOL_F32)}



# ------------------------------------------------------------------------------------------------
# training-step kernels: each against torch autograd of the same op (fp32 / fp64 on the GPU box)
# ------------------------------------------------------------------------------------------------
TOL_GRAD = 1.5e-2   # bf16 gradients in, bf16 gradients out


def check_layernorm_bwd():
    ops = _ops()
    out = {}
    dim = 728
    for rows, xdt in ((37, torch.float32), (3 * 7 * 50, torch.float32), (1000, torch.bfloat16)):
        x = (_rand(rows, dim, seed=rows) * 2 + 0.3).to(xdt)
        g = _rand(dim, seed=1) * 0.2 + 1
        b = _rand(dim, seed=2) * 0.1
        dy = _rand(rows, dim, seed=3).to(torch.bfloat16)
        xr = x.float().clone().requires_grad_(True)
        gr = g.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        F.layer_norm(xr, (dim,), gr, br, 1e-5).backward(dy.float())
        dg, db = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
        dx = ops.layernorm_bwd(dy, x, g, dg, db)
        out[f"dx_{rows}"] = _assert_close("ln_bwd dx", dx, xr.grad, TOL_BF16)
        out[f"dg_{rows}"] = _assert_close("ln_bwd dgamma", dg, gr.grad, 2e-4)
        out[f"db_{rows}"] = _assert_close("ln_bwd dbeta", db, br.grad, 2e-4)
        # accumulate mode: g += dx, bf16 copy
        acc = _rand(rows, dim, seed=9)
        want = acc + xr.grad
        gbf = torch.empty(rows, dim, dtype=torch.bfloat16, device=DEV)
        dg.zero_(); db.zero_()
        assert ops.layernorm_bwd(dy, x, g, dg, db, g_accum=acc, g_bf16=gbf) is None
        out[f"acc_{rows}"] = _assert_close("ln_bwd accumulate", acc, want, 5e-3)
        out[f"accbf_{rows}"] = _assert_close("ln_bwd accumulate bf16 copy", gbf, want, TOL_BF16)
        # fused bias gradient: column sums of the rows produced (dx, or the updated g), added to what is there; row-pitched
        # operands (the training step's layout)
        cs = torch.ones(dim, device=DEV)
        dyp, dxw = ops.pad_rows(dy), xr.grad
        dg.zero_(); db.zero_()
        dxp = ops.layernorm_bwd(dyp, x, g, dg, db, out_colsum=cs)
        assert dxp.stride(0) == ops.row_pitch(dim) and torch.equal(dxp, dx), "pitched operands must not change dx"
        out[f"cs_dx_{rows}"] = _assert_close("ln_bwd column sums of dx", cs, 1 + dxw.sum(0), 5e-4)
        acc2, cs2 = _rand(rows, dim, seed=9), torch.zeros(dim, device=DEV)
        gbp = ops.empty_rows((rows, dim), torch.bfloat16, DEV)
        ops.layernorm_bwd(dyp, x, g, dg, db, g_accum=acc2, g_bf16=gbp, out_colsum=cs2)
        assert torch.equal(acc2, acc) and torch.equal(gbp, gbf)
        out[f"cs_acc_{rows}"] = _assert_close("ln_bwd column sums of g", cs2, want.sum(0), 5e-4)
    # fused self-subtract backward (module.py:192): x [B, F, P, D]
    bsz, f, p = 2, 7, 11
    x = _rand(bsz, f, p, dim, seed=5) * 1.5
    g = _rand(dim, seed=1) * 0.2 + 1
    b = _rand(dim, seed=2) * 0.1
    dxn = _rand(bsz, f, p, dim, seed=6).to(torch.bfloat16)
    dres = _rand(bsz, f, p, dim, seed=7).to(torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    xn = F.layer_norm(xr, (dim,), gr, br, 1e-5)
    res = torch.cat((xn[:, :2], xn[:, 2:] - xn[:, 1:-1]), dim=1)
    (xn * dxn.float()).sum().add((res * dres.float()).sum()).backward()
    dg, db = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    acc = torch.zeros_like(x)
    ops.layernorm_bwd(dxn, x, g, dg, db, g_accum=acc, dy2=dres, frames=f, tokens_per_frame=p)
    out["diff_dx"] = _assert_close("ln_bwd + self-subtract dx", acc, xr.grad, 5e-3)
    out["diff_dg"] = _assert_close("ln_bwd + self-subtract dgamma", dg, gr.grad, 2e-4)
    out["diff_db"] = _assert_close("ln_bwd + self-subtract dbeta", db, br.grad, 2e-4)
    return out


def check_gelu_cast_transpose():
    ops = _ops()
    out = {}
    x = (_rand(1000, 2912, seed=1) * 2).to(torch.bfloat16)
    dy = _rand(1000, 2912, seed=2).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    y = F.gelu(xr)
    y.backward(dy.float())
    out["gelu"] = _assert_close("gelu fwd", ops.gelu(x), y, TOL_BF16)
    out["gelu_bwd"] = _assert_close("gelu bwd", ops.gelu_bwd(dy, x), xr.grad, TOL_BF16)
    for rows in (1000, 37):      # with the fused column sums (bias gradient of net[0]); several / one row per CTA slot
        cs = torch.ones(2912, device=DEV)
        got = ops.gelu_bwd(dy[:rows], x[:rows], colsum=cs)
        assert torch.equal(got, ops.gelu_bwd(dy[:rows], x[:rows])), "the column sums must not change dx"
        out[f"gelu_bwd_colsum_{rows}"] = _assert_close("gelu bwd column sums", cs, 1 + xr.grad[:rows].sum(0), 1e-3)
    xs = (_rand(300, 64, seed=4) * 2).to(torch.bfloat16)      # narrow matrix: many rows per CTA
    dys = _rand(300, 64, seed=5).to(torch.bfloat16)
    xsr = xs.float().requires_grad_(True)
    F.gelu(xsr).backward(dys.float())
    cs = torch.zeros(64, device=DEV)
    out["gelu_bwd_narrow"] = _assert_close("gelu bwd narrow", ops.gelu_bwd(dys, xs, colsum=cs), xsr.grad, TOL_BF16)
    out["gelu_bwd_narrow_cs"] = _assert_close("gelu bwd narrow column sums", cs, xsr.grad.sum(0), 1e-3)
    f = _rand(333, 728, seed=3)
    out["cast"] = _assert_close("cast", ops.cast_bf16(f), f, TOL_BF16)
    for (m, c) in ((5068, 728), (64, 64), (1001, 2912), (77, 8)):
        a = _rand(m, c, seed=m).to(torch.bfloat16)
        cs = torch.ones(c, device=DEV)
        t = ops.transpose(a, cs)
        ld = (m + 7) // 8 * 8
        assert tuple(t.shape) == (c, ld)
        assert torch.equal(t[:, :m], a.t()), f"transpose {m}x{c}"
        assert ld == m or float(t[:, m:].abs().max()) == 0.0, "transpose pad must be zero"
        out[f"colsum_{m}"] = _assert_close("colsum", cs, 1 + a.float().sum(0), 1e-4)
    return out


def check_gemm_wgrad():
    ops = _ops()
    _noTF32()
    out = {}
    for (rows, n, k) in ((5068, 1024, 728), (20000, 2912, 728), (5068, 728, 2912), (9000, 128, 64), (700, 512, 728),
                         (40, 728, 512)):
        dy = _rand(rows, n, seed=n).to(torch.bfloat16)
        x = _rand(rows, k, seed=k + 1).to(torch.bfloat16)
        dw = _rand(n, k, seed=3)                     # accumulates on top of existing gradient content
        want = dw.double() + dy.double().t() @ x.double()
        ops.gemm_wgrad(ops.transpose(dy), ops.transpose(x), rows, dw)
        out[f"{rows}x{n}x{k}"] = _assert_close(f"wgrad {rows}x{n}x{k}", dw, want.float(), 2e-4)
        # in-place (MN-major operand) weight gradient + bias gradient
        dw2 = _rand(n, k, seed=3)
        db = torch.ones(n, device=DEV)
        ops.wgrad(dy, x, dw2, bias_grad=db)
        out[f"mn_{rows}x{n}x{k}"] = _assert_close(f"wgrad MN-major {rows}x{n}x{k}", dw2, want.float(), 2e-4)
        out[f"db_{rows}x{n}"] = _assert_close(f"bias grad {rows}x{n}", db, 1 + dy.float().sum(0), 2e-4)
    return out


def check_attn_temporal_bwd():
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    # f <= 8: registers-only kernel; 9 <= f <= 48: shared-memory / ldmatrix kernel (2 or 3 frame tiles)
    for (b, f, p) in ((2, 7, 362), (1, 2, 5), (3, 8, 33), (1, 33, 362), (2, 9, 21), (1, 16, 17), (1, 17, 30), (1, 32, 9),
                      (1, 48, 5)):
        rows = b * f * p
        qk = (_rand(rows, 1024, seed=f) * 1.5).to(torch.bfloat16)
        v = _rand(rows, 512, seed=f + 1).to(torch.bfloat16)
        do = _rand(rows, 512, seed=f + 2).to(torch.bfloat16)
        qkr = qk.double().requires_grad_(True)
        vr = v.double().requires_grad_(True)
        split = lambda t: t.reshape(b, f, p, heads, 64).permute(0, 3, 2, 1, 4)   # b h p f d
        q, k = split(qkr[:, :512]), split(qkr[:, 512:])
        a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
        o = torch.matmul(a, split(vr)).permute(0, 3, 2, 1, 4).reshape(rows, 512)
        o.backward(do.double())
        dqk, dv = ops.attn_temporal_bwd(qk, v, do, b, f, p, heads, scale)
        out[f"dqk_{f}"] = _assert_close("attn_t bwd dqk", dqk, qkr.grad.float(), TOL_GRAD)
        out[f"dv_{f}"] = _assert_close("attn_t bwd dv", dv, vr.grad.float(), TOL_GRAD)
    return out


def check_attn_spatial_bwd():
    ops = _ops()
    out = {}
    heads, scale = 8, 0.125
    for (bf, p, amp) in ((2, 362, 1.5), (3, 130, 1.0), (1, 50, 1.0), (21, 384, 1.0)):
        rows = bf * p
        qkv = (_rand(rows, 1536, seed=p + bf) * amp).to(torch.bfloat16)
        do = _rand(rows, 512, seed=p + 1).to(torch.bfloat16)
        r = qkv.double().requires_grad_(True)
        split = lambda t: t.reshape(bf, p, heads, 64).permute(0, 2, 1, 3)
        q, k, v = split(r[:, :512]), split(r[:, 512:1024]), split(r[:, 1024:])
        a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
        o_ref = torch.matmul(a, v).permute(0, 2, 1, 3).reshape(rows, 512)
        o_ref.backward(do.double())
        o, lse = ops.attn_spatial_lse(qkv, bf, p, heads, scale)
        name = f"{bf}x{p}"
        out[f"fwd_{name}"] = _assert_close(f"attn_s lse-mode out {name}", o, o_ref.float(), 1.5e-2)
        lse_ref = torch.logsumexp(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1) / math.log(2.0)
        out[f"lse_{name}"] = _assert_close(f"attn_s lse {name}", lse, lse_ref.float(), 2e-3)
        dqkv = ops.attn_spatial_bwd(qkv, o, do, lse, bf, p, heads, scale)
        for nm, sl in (("dq", slice(0, 512)), ("dk", slice(512, 1024)), ("dv", slice(1024, 1536))):
            out[f"{nm}_{name}"] = _assert_close(f"attn_s bwd {nm} {name}", dqkv[:, sl], r.grad[:, sl].float(), TOL_GRAD)
    return out


def check_head_token_bwd():
    ops = _ops()
    out = {}
    b, f, p, d = 3, 7, 362, 728
    tokens = _rand(b, f, p, d, seed=1) * 2
    prm = [(_rand(d, seed=2) * 0.2 + 1), _rand(d, seed=3) * 0.1, (_rand(d, seed=4) * 0.2 + 1), _rand(d, seed=5) * 0.1,
           _rand(d, seed=6) * 0.05, _rand(1, seed=7)]
    dz = _rand(b, seed=8)
    tr = tokens.clone().requires_grad_(True)
    pr = [t.clone().requires_grad_(True) for t in prm]
    x = F.layer_norm(tr[:, 0, 0], (d,), pr[0], pr[1], 1e-5)
    z = F.layer_norm(x, (d,), pr[2], pr[3], 1e-5) @ pr[4][:, None] + pr[5]
    z.backward(dz[:, None])
    g = torch.zeros_like(tokens)
    grads = [torch.zeros_like(t) for t in prm]
    ops.head_bwd(tokens, dz, prm[0], prm[1], prm[2], prm[3], prm[4], g, *grads)
    out["g"] = _assert_close("head_bwd dtokens", g, tr.grad, 1e-4)
    for i, nm in enumerate(("norm_g", "norm_b", "head_g", "head_b", "head_w", "head_bias")):
        out[nm] = _assert_close(f"head_bwd d{nm}", grads[i], pr[i].grad, 1e-4)
    # token-build backward
    gt = _rand(b, f, p, d, seed=11)
    dpos, dsp, dtm = torch.zeros(f - 1, p, d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.token_bwd(gt, dpos, dsp, dtm)
    out["dpos"] = _assert_close("token_bwd dpos", dpos, gt[:, 1:].sum(0), 1e-5)
    out["dspace"] = _assert_close("token_bwd dspace", dsp, gt[:, 1:, 0].sum((0, 1)), 1e-5)
    out["dtemporal"] = _assert_close("token_bwd dtemporal", dtm, gt[:, 0].sum((0, 1)), 1e-5)
    return out


def check_adamw():
    ops = _ops()
    out = {}
    n = 4 * 1001
    p0 = _rand(n, seed=1)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = _rand(n, seed=10 + step)
        ref.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g * 4.0, m, v, 3e-3, (0.9, 0.999), 1e-8, 0.05, step, grad_scale=0.25)
        out[f"step{step}"] = _assert_close(f"adamw step {step}", p, ref.data, 1e-5)
    return out



def check_entry_train_kernels():
    """BatchNorm (batch statistics) fwd/bwd, pool+add with arg-max and its backward, depthwise weight gradient,
    block-input gradient, im2col^T operands — each against torch autograd."""
    ops = _ops()
    out = {}
    # ---- BatchNorm2d training mode ----
    for (n, h, w, c, relu) in ((3, 19, 17, 728, False), (2, 37, 37, 64, True), (2, 21, 20, 32, True)):
        x = (_rand(n, h, w, c, seed=c) * 1.7 + 0.4).to(torch.bfloat16)
        bn = torch.nn.BatchNorm2d(c).to(DEV)
        with torch.no_grad():
            bn.weight.copy_(_rand(c, seed=1) * 0.2 + 1); bn.bias.copy_(_rand(c, seed=2) * 0.1)
            bn.running_mean.copy_(_rand(c, seed=3) * 0.1); bn.running_var.copy_(_rand(c, seed=4).abs() + 0.5)
        ref_bn = torch.nn.BatchNorm2d(c).to(DEV)
        ref_bn.load_state_dict(bn.state_dict())
        xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        yr = ref_bn(xr)
        yr = torch.relu(yr) if relu else yr
        dy = _rand(n, h, w, c, seed=9).to(torch.bfloat16)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        y, st = ops.batchnorm_train(x, bn, relu)
        tag = f"bn{c}"
        out[tag + "_y"] = _assert_close("bn train y", y.float(), yr.permute(0, 2, 3, 1), TOL_BF16)
        out[tag + "_rm"] = _assert_close("bn running_mean", bn.running_mean, ref_bn.running_mean, 1e-4)
        out[tag + "_rv"] = _assert_close("bn running_var", bn.running_var, ref_bn.running_var, 1e-4)
        assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 1
        dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        dx = ops.batchnorm_bwd(dy, x, st, dg, db, relu)
        out[tag + "_dx"] = _assert_close("bn bwd dx", dx.float(), xr.grad.permute(0, 2, 3, 1), 1.5e-2)
        out[tag + "_dg"] = _assert_close("bn bwd dgamma", dg, ref_bn.weight.grad, 5e-3)
        out[tag + "_db"] = _assert_close("bn bwd dbeta", db, ref_bn.bias.grad, 5e-3)
    # ---- maxpool(3,2,1) + skip with arg-max, and its backward ----
    for (n, h, w, c) in ((2, 37, 37, 64), (3, 10, 7, 8)):
        # distinct values per window (no bf16 ties): arg-max routing is then unambiguous
        x = (torch.randperm(n * h * w * c, generator=torch.Generator().manual_seed(5)).float().reshape(n, h, w, c)
             / (n * h * w * c) * 200 - 100).to(DEV)
        xb = x.to(torch.bfloat16)
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        skip = _rand(n, ho, wo, c, seed=6).to(torch.bfloat16)
        xr = xb.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        yr = F.max_pool2d(xr, 3, 2, 1) + skip.float().permute(0, 3, 1, 2)
        dy = _rand(n, ho, wo, c, seed=7).to(torch.bfloat16)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        y, amax = ops.pool_add_idx(xb, skip)
        out[f"pool_y_{h}"] = _assert_close("pool_add_idx y", y.float(), yr.permute(0, 2, 3, 1), TOL_BF16)
        dx = ops.pool_bwd(dy, amax, h, w)
        out[f"pool_dx_{h}"] = _assert_close("pool_bwd dx", dx.float(), xr.grad.permute(0, 2, 3, 1), TOL_BF16)
    # token variant + token-gradient gather
    b, t, c = 2, 3, 16
    xb = _rand(b * t, 37, 37, c, seed=8).to(torch.bfloat16)
    skip = _rand(b * t, 19, 19, c, seed=9).to(torch.bfloat16)
    pos = _rand(t, 362, c, seed=10)
    tokens = torch.zeros(b, t + 1, 362, c, device=DEV)
    _, amax = ops.pool_add_idx(xb, skip, tokens=tokens, pos_emb=pos, t_frames=t)
    ref = (F.max_pool2d(xb.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1) + skip.float()).reshape(b, t, 361, c)
    out["pool_tokens"] = _assert_close("pool_add_idx tokens", tokens[:, 1:, 1:], ref + pos[None, :, 1:], 1e-5)
    gt = _rand(b, t + 1, 362, c, seed=11)
    out["token_gather"] = _assert_close("token_grad_gather", ops.token_grad_gather(gt).float(),
                                        gt[:, 1:, 1:].reshape(b * t, 19, 19, c), TOL_BF16)
    # ---- depthwise weight gradient + data gradient via flipped taps, block-input gradient ----
    for (n, h, w, c, relu_in) in ((2, 37, 37, 64, True), (3, 19, 21, 728, False), (1, 9, 9, 8, True)):
        xb = _rand(n, h, w, c, seed=c).to(torch.bfloat16)
        wt = _rand(3, 3, c, seed=2) * 0.3
        dy = _rand(n, h, w, c, seed=3).to(torch.bfloat16)
        xr = xb.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        wr = wt.permute(2, 0, 1).unsqueeze(1).clone().requires_grad_(True)           # [c, 1, 3, 3]
        yr = F.conv2d(torch.relu(xr) if relu_in else xr, wr, None, 1, 1, 1, groups=c)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        dw = torch.zeros(3, 3, c, device=DEV)
        ops.dwconv3x3_wgrad(xb, dy, dw, relu_in)
        out[f"dw_wgrad_{c}"] = _assert_close("dwconv wgrad", dw, wr.grad[:, 0].permute(1, 2, 0), 2e-3)
        d_main = ops.dwconv3x3(dy, wt.flip(0, 1).contiguous(), relu_in=False)
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        d_skip = _rand(n, ho, wo, c, seed=4).to(torch.bfloat16)
        dx = ops.block_input_grad(d_main, xb, d_skip, relu_in)
        want = xr.grad.permute(0, 2, 3, 1).clone()
        want[:, ::2, ::2] += d_skip.float()
        out[f"block_in_{c}"] = _assert_close("dwconv dgrad + block input grad", dx.float(), want, 1.5e-2)
    # ---- im2col^T operands ----
    xb = _rand(2, 11, 9, 32, seed=1).to(torch.bfloat16)
    cols = ops.im2col_t(xb)
    m = 2 * 9 * 7
    ref = torch.stack([xb[:, ky:ky + 9, kx:kx + 7, :].reshape(m, 32) for ky in range(3) for kx in range(3)], 0)  # [9, m, 32]
    assert torch.equal(cols[:, :m].reshape(9, 32, m), ref.permute(0, 2, 1)), "im2col_t"
    # several 256-pixel tiles, image boundaries inside a tile, a ragged last tile, zero fill past the last pixel
    for (nn, hh, ww, cc) in ((3, 13, 23, 32), (2, 19, 18, 64)):
        xb = _rand(nn, hh, ww, cc, seed=hh).to(torch.bfloat16)
        cols = ops.im2col_t(xb)
        m = nn * (hh - 2) * (ww - 2)
        ref = torch.stack([xb[:, ky:ky + hh - 2, kx:kx + ww - 2, :].reshape(m, cc) for ky in range(3) for kx in range(3)], 0)
        assert torch.equal(cols[:, :m].reshape(9, cc, m), ref.permute(0, 2, 1)), f"im2col_t {nn}x{hh}x{ww}x{cc}"
        assert not cols[:, m:].any(), "im2col_t: columns past the last pixel must be zero"
    xf = _rand(2, 3, 21, 19, seed=2)
    cols = ops.im2col_t_stem(xf)
    ho, wo = 10, 9
    unf = F.unfold(xf, 3, stride=2)                                 # [2, 27, ho*wo], k = ci*9 + ky*3 + kx
    ref = unf.permute(1, 0, 2).reshape(27, 2 * ho * wo)
    out["im2col_stem"] = _assert_close("im2col_t_stem", cols[:27, :2 * ho * wo].float(), ref, TOL_BF16)
    assert float(cols[27:].abs().max()) == 0.0
    # conv stem without ReLU
    wt = _rand(32, 3, 3, 3, seed=3) * 0.2
    y = ops.conv_stem_raw(xf, wt)
    out["stem_raw"] = _assert_close("conv_stem_raw", y.float(), F.conv2d(xf, wt, None, 2).permute(0, 2, 3, 1), TOL_BF16)
    return out


CHECKS = {
    "layernorm": check_layernorm,
    "layernorm_diff": check_layernorm_diff,
    "gemm_basic": check_gemm_basic,
    "gemm_shapes": check_gemm_shapes,
    "gemm_lnfold": check_gemm_lnfold,
    "gemm_mlp_fusions": check_gemm_mlp_fusions,
    "gemm_f32": check_gemm_f32,
    "conv3x3": check_conv3x3,
    "conv_stem": check_conv_stem,
    "conv_stem_u8": check_conv_stem_u8,
    "attn_spatial_spiky": check_attn_spatial_spiky,
    "xception_tail": check_xception_tail,
    "dwconv": check_dwconv,
    "sepconv_fused": check_sepconv_fused,
    "pool_subsample_tokens": check_pool_subsample_tokens,
    "attn_temporal": check_attn_temporal,
    "attn_spatial_f32": check_attn_spatial_f32,
    "attn_spatial_bf16": check_attn_spatial_bf16,
    "attn_joint": check_attn_joint,
    "token_build": check_token_build,
    "head": check_head,
    "layernorm_bwd": check_layernorm_bwd,
    "gelu_cast_transpose": check_gelu_cast_transpose,
    "gemm_wgrad": check_gemm_wgrad,
    "attn_temporal_bwd": check_attn_temporal_bwd,
    "attn_spatial_bwd": check_attn_spatial_bwd,
    "head_token_bwd": check_head_token_bwd,
    "adamw": check_adamw,
    "entry_train_kernels": check_entry_train_kernels,
}
