"""Tensor-level wrappers over the C ABI: one Python function per kernel family.

torch is used here for device memory, streams and nothing else — every function validates its
arguments, allocates the output with `torch.empty`, and hands raw device pointers plus the current CUDA
stream to libistvt_b200.so.  Non-CUDA tensors are rejected (there is no CPU path).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
_ROW_PITCH = os.environ.get("ISTVT_ROW_PITCH", "1") != "0"      # 0: every matrix dense (A/B measurements)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"istvt_b200: unsupported dtype {t.dtype} (bf16 or fp32 only)")


def _torch_dt(code: int) -> torch.dtype:
    return torch.bfloat16 if code == BF16 else torch.float32


def _chk(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("istvt_b200: tensors must live on a CUDA device (no CPU fallback by design)")
        if not t.is_contiguous():
            raise ValueError("istvt_b200: tensors must be contiguous")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("istvt_b200: all tensors of one call must be on the same device")
    assert dev is not None
    return dev


def _ld(t: torch.Tensor) -> int:
    """Row pitch (elements) of a row-pitched matrix: last axis dense, all leading axes collapse to one row index with a
    uniform pitch >= shape[-1] (a contiguous tensor, or a `[..., :dim]` view of one whose rows are padded)."""
    if t.dim() < 2:
        return t.shape[-1]
    if t.stride(-1) != 1 or t.stride(-2) < t.shape[-1]:
        raise ValueError("istvt_b200: the last axis must be dense")
    ld = t.stride(-2)
    step = ld
    for i in range(t.dim() - 2, 0, -1):
        step *= t.shape[i]
        if t.shape[i - 1] > 1 and t.stride(i - 1) != step:
            raise ValueError("istvt_b200: rows must have a uniform pitch")
    return ld


def _chk_rows(*tensors: Optional[torch.Tensor]) -> torch.device:
    """_chk for row-pitched matrices (see _ld) instead of contiguous tensors."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("istvt_b200: tensors must live on a CUDA device (no CPU fallback by design)")
        _ld(t)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("istvt_b200: all tensors of one call must be on the same device")
    assert dev is not None
    return dev


def row_pitch(dim: int, dtype: torch.dtype = torch.bfloat16) -> int:
    """Pitch (elements) that makes every row of a [rows, dim] matrix start on a 128-byte line: 728 bf16 -> 768.
    Rows that already start on 64-byte boundaries are left alone (dim 2912: a 128-byte box row then spans two lines
    on odd rows but the same four 32-byte sectors; padding to 2944 measured +1 % on A and -3 % on W, r6n)."""
    size = torch.empty((), dtype=dtype).element_size()
    if dim * size % 64 == 0 or not _ROW_PITCH:
        return dim
    per_line = 128 // size
    return (dim + per_line - 1) // per_line * per_line


def empty_rows(shape, dtype: torch.dtype, device) -> torch.Tensor:
    """torch.empty(shape) whose rows (last axis) sit at the 128-byte aligned pitch row_pitch(shape[-1]): returned as
    the `[..., :dim]` view.  TMA box rows of such an operand never straddle two lines (profiles/README.md r6n)."""
    dim = shape[-1]
    ld = row_pitch(dim, dtype)
    full = torch.empty(*shape[:-1], ld, dtype=dtype, device=device)
    return full if ld == dim else full[..., :dim]


def pad_rows(t: torch.Tensor) -> torch.Tensor:
    """Copy of a [rows, dim] matrix at the aligned pitch (weights, at pack time)."""
    out = empty_rows(tuple(t.shape), t.dtype, t.device)
    out.copy_(t)
    return out


def f32_aligned(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, 16-byte aligned view or copy of a parameter.  Parameters of an `nn.DataParallel` replica are
    views into one coalesced broadcast buffer (torch's `broadcast_coalesced` packs them back to back), so their start
    is only element-aligned; the kernels read vectors of 16 bytes."""
    t = t.detach().float().contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


# ----------------------------------------------------------------------------------------------
# Launch-level profiling (bench.py's roofline): when a recorder is installed every wrapper brackets its
# kernel launch with a CUDA event pair recorded on the launching stream and notes the ALGORITHMIC work
# of that launch (flops for tensor-bound kernels, bytes for HBM-bound ones; DESIGN.md §4 has the formulas).
# ----------------------------------------------------------------------------------------------
class LaunchRecorder:
    def __init__(self):
        self.records = []   # (family, flops, bytes, start_event, end_event)

    def summary(self):
        """{family: {"launches", "ms", "flops", "bytes"}} — call after torch.cuda.synchronize()."""
        out = {}
        for fam, fl, by, e0, e1 in self.records:
            d = out.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        return out


_recorder: Optional[LaunchRecorder] = None


def set_recorder(rec: Optional[LaunchRecorder]) -> None:
    global _recorder
    _recorder = rec


class _launch:
    """`with _launch(dev, family, flops, bytes): <one C-ABI call>` — device guard + optional event pair."""

    __slots__ = ("dev", "prev", "fam", "flops", "bytes", "e0")

    def __init__(self, dev: torch.device, fam: str, flops: float = 0.0, nbytes: float = 0.0):
        self.dev, self.fam, self.flops, self.bytes = dev, fam, flops, nbytes
        self.prev = None
        self.e0 = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev.index:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        if _recorder is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.dev))
            if _recorder is not None:
                _recorder.records.append((self.fam, self.flops, self.bytes, self.e0, e1))
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def _nbytes(*ts: Optional[torch.Tensor]) -> int:
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


# ----------------------------------------------------------------------------------------------
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_dtype: torch.dtype,
              eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _chk(gamma, beta)
    _chk_rows(x, out)
    dim = x.shape[-1]
    rows = x.numel() // dim
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=dev)
    with _launch(dev, "layernorm", 0.0, _nbytes(x, out)):
        _lib.check(_lib.lib().istvt_layernorm_fwd_ld(_ptr(x), _dt(x), _ld(x), _ptr(gamma), _ptr(beta), _ptr(out),
                                                     _dt(out), _ld(out), rows, dim, eps, _stream(dev)),
                   "istvt_layernorm_fwd_ld")
    return out


def layernorm_diff(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_dtype: torch.dtype,
                   eps: float = 1e-5, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x: fp32 [B, F, P, D] -> (LN(x), self-subtract difference), both [B, F, P, D] in out_dtype."""
    dev = _chk(gamma, beta)
    _chk_rows(x)
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError("layernorm_diff expects an fp32 [B, F, P, D] token tensor")
    b, f, p, d = x.shape
    if out is None:
        xn = torch.empty(x.shape, dtype=out_dtype, device=dev)
        diff = torch.empty(x.shape, dtype=out_dtype, device=dev)
    else:
        xn, diff = out
        _chk_rows(xn, diff)
        if _ld(xn) != _ld(diff):
            raise ValueError("layernorm_diff: xn and diff must have the same row pitch")
    with _launch(dev, "layernorm_diff", 0.0, _nbytes(x, xn, diff)):
        _lib.check(_lib.lib().istvt_layernorm_diff_fwd_ld(_ptr(x), _ld(x), _ptr(gamma), _ptr(beta), _ptr(xn), _ptr(diff),
                                                          _dt(xn), _ld(xn), b, f, p, d, eps, _stream(dev)),
                   "istvt_layernorm_diff_fwd_ld")
    return xn, diff


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[m, n] = act(a[m, :] . w[n, :] + bias[n]) + residual[m, n].  a: [..., K], w: [N, K].

    bf16 operands run the tcgen05 kernel, fp32 operands the SIMT validation kernel.
    `residual` must be fp32; pass `out=residual` for the in-place residual update.
    """
    dev = _chk(bias) if bias is not None else a.device
    _chk_rows(a, w, residual, out)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if w.shape[1] != k or a.dtype != w.dtype:
        raise ValueError(f"gemm: a {tuple(a.shape)} {a.dtype} vs w {tuple(w.shape)} {w.dtype}")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != n):
        raise ValueError("gemm: bias must be fp32 [N]")
    if residual is not None and (residual.dtype != torch.float32 or residual.numel() != m * n):
        raise ValueError("gemm: residual must be fp32 [M, N]")
    if out is None:
        if out_dtype is None:
            out_dtype = torch.float32 if (residual is not None or a.dtype == torch.float32) else a.dtype
        out = torch.empty(*a.shape[:-1], n, dtype=out_dtype, device=dev)
    elif out.numel() != m * n:
        raise ValueError("gemm: out has the wrong size")
    st = _stream(dev)
    with _launch(dev, "gemm_bf16" if a.dtype == torch.bfloat16 else "gemm_f32", 2.0 * m * n * k,
                 _nbytes(a, w, out, residual)):
        if a.dtype == torch.bfloat16:
            _lib.check(_lib.lib().istvt_gemm_fwd(_ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(out), _ld(out), _dt(out), m, n, k,
                                                 _ptr(bias), _ptr(residual), _ld(residual) if residual is not None else n,
                                                 act, st), "istvt_gemm_fwd")
        else:
            if out.dtype != torch.float32:
                raise ValueError("gemm: fp32 operands need an fp32 output")
            _lib.check(_lib.lib().istvt_gemm_f32_fwd(_ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(out), _ld(out), m, n, k,
                                                     _ptr(bias), _ptr(residual),
                                                     _ld(residual) if residual is not None else n, act, st),
                       "istvt_gemm_f32_fwd")
    return out


def gemm_act_dual(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_GELU
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(act(a . w^T + bias), a . w^T + bias), both bf16 [rows, N], from one accumulator (istvt_gemm_act_dual_fwd): the
    training forward of FeedForward.net[0:2] keeps the pre-activation for the backward without a GELU pass."""
    dev = _chk(bias) if bias is not None else a.device
    _chk_rows(a, w)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or w.shape[1] != k:
        raise ValueError("gemm_act_dual: bf16 a [M, K] and w [N, K]")
    out = torch.empty(*a.shape[:-1], n, dtype=torch.bfloat16, device=dev)
    pre = torch.empty(*a.shape[:-1], n, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "gemm_bf16", 2.0 * m * n * k, _nbytes(a, w, out, pre)):
        _lib.check(_lib.lib().istvt_gemm_act_dual_fwd(_ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(out), n, _ptr(pre), n, m, n, k,
                                                      _ptr(bias), act, _stream(dev)), "istvt_gemm_act_dual_fwd")
    return out, pre


def gemm_dgelu(a: torch.Tensor, w: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    """(a . w^T) o gelu'(pre), bf16 [rows, N] (istvt_gemm_dgelu_fwd): the data gradient through FeedForward.net[3] and
    the GELU in one kernel."""
    dev = a.device
    _chk_rows(a, w, pre)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or w.shape[1] != k or pre.dtype != torch.bfloat16 or \
            pre.numel() != m * n:
        raise ValueError("gemm_dgelu: bf16 a [M, K], w [N, K], pre [M, N]")
    out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "gemm_bf16", 2.0 * m * n * k, _nbytes(a, w, out, pre)):
        _lib.check(_lib.lib().istvt_gemm_dgelu_fwd(_ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(out), n, _ptr(pre), _ld(pre), m, n,
                                                   k, _stream(dev)), "istvt_gemm_dgelu_fwd")
    return out


def gemm_rowstats(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], row_stats: torch.Tensor) -> torch.Tensor:
    """out = a . w^T + bias (bf16) and row_stats[m, g] = (sum, sum of squared deviations from the group mean) of every
    64-column group of the output row — the producer half of the LayerNorm fold (istvt_gemm_rowstats_fwd).
    row_stats: fp32 [M, ceil(N / 64), 2], fully overwritten."""
    dev = _chk(bias, row_stats)
    _chk_rows(a, w)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or w.shape[1] != k:
        raise ValueError("gemm_rowstats: bf16 a [M, K] and w [N, K]")
    if row_stats.dtype != torch.float32 or row_stats.numel() != 2 * m * ((n + 63) // 64):
        raise ValueError("gemm_rowstats: row_stats must be fp32 [M, ceil(N / 64), 2]")
    out = empty_rows((*a.shape[:-1], n), torch.bfloat16, dev)      # the next GEMM's A operand: aligned pitch
    with _launch(dev, "gemm_bf16", 2.0 * m * n * k, _nbytes(a, w, out)):
        _lib.check(_lib.lib().istvt_gemm_rowstats_fwd(_ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(out), _ld(out), m, n, k,
                                                      _ptr(bias), _ptr(row_stats), _stream(dev)),
                   "istvt_gemm_rowstats_fwd")
    return out


def ln_stats_finalize(row_stats: torch.Tensor, dim: int, eps: float = 1e-5) -> torch.Tensor:
    """(mean, rstd) per row, fp32 [M, 2], from gemm_rowstats' per-group partials (istvt_ln_stats_finalize)."""
    dev = _chk(row_stats)
    groups = (dim + 63) // 64
    if row_stats.dtype != torch.float32 or row_stats.numel() % (2 * groups):
        raise ValueError("ln_stats_finalize: row_stats must be fp32 [M, ceil(dim / 64), 2]")
    m = row_stats.numel() // (2 * groups)
    out = torch.empty(m, 2, dtype=torch.float32, device=dev)
    with _launch(dev, "layernorm", 0.0, _nbytes(row_stats, out)):
        _lib.check(_lib.lib().istvt_ln_stats_finalize(_ptr(row_stats), _ptr(out), m, dim, eps, _stream(dev)),
                   "istvt_ln_stats_finalize")
    return out


def gemm_lnfold(a: torch.Tensor, w_folded: torch.Tensor, mu_rstd: torch.Tensor, w_rowsum: torch.Tensor,
                shift: torch.Tensor) -> torch.Tensor:
    """LayerNorm(a) . W^T without the LayerNorm pass — the consumer half of the fold (istvt_gemm_lnfold_fwd):
    out = rstd (a . w_folded^T - mu w_rowsum) + shift with (mu, rstd) per row from ln_stats_finalize."""
    dev = _chk(mu_rstd, w_rowsum, shift)
    _chk_rows(a, w_folded)
    k = a.shape[-1]
    m = a.numel() // k
    n = w_folded.shape[0]
    if a.dtype != torch.bfloat16 or w_folded.dtype != torch.bfloat16 or w_folded.shape[1] != k:
        raise ValueError("gemm_lnfold: bf16 a [M, K] and w_folded [N, K]")
    if mu_rstd.numel() != 2 * m or w_rowsum.numel() != n or shift.numel() != n or \
            any(t.dtype != torch.float32 for t in (mu_rstd, w_rowsum, shift)):
        raise ValueError("gemm_lnfold: mu_rstd fp32 [M, 2], w_rowsum / shift fp32 [N]")
    out = torch.empty(*a.shape[:-1], n, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "gemm_bf16", 2.0 * m * n * k, _nbytes(a, w_folded, out)):
        _lib.check(_lib.lib().istvt_gemm_lnfold_fwd(_ptr(a), _ld(a), _ptr(w_folded), _ld(w_folded), _ptr(out), n, m, n, k,
                                                    _ptr(mu_rstd), _ptr(w_rowsum), _ptr(shift), _stream(dev)),
                   "istvt_gemm_lnfold_fwd")
    return out


_CONV2_PAIR = os.environ.get("ISTVT_CONV2_KERNEL", "pair") == "pair"   # taps | strip | gather: istvt_conv3x3_fwd's own switch


def conv3x3_pair_weights(w: torch.Tensor, w_in: int) -> torch.Tensor:
    """conv2's weights [64, 3, 3, 32] (bf16) rearranged for the pixel-pair kernel: [128, taps * 64], taps = 7 / 6 for an odd
    / even input width.  Memoised ON the weight tensor (per parity, invalidated by its version counter): an inference pack
    is rearranged once, a training step — whose folded weight is a new tensor every step — once per step."""
    memo = getattr(w, "_istvt_pair", None)
    if memo is None:
        memo = {}
        w._istvt_pair = memo
    hit = memo.get(w_in & 1)
    if hit is None or hit[0] != w._version:
        taps = 7 if (w_in & 1) else 6
        wp = torch.empty(128, taps * 64, dtype=torch.bfloat16, device=w.device)
        with _launch(w.device, "conv3x3_pack", 0.0, _nbytes(w, wp)):
            _lib.check(_lib.lib().istvt_conv3x3_pair_pack(_ptr(w), _ptr(wp), w_in, _stream(w.device)),
                       "istvt_conv3x3_pair_pack")
        hit = (w._version, wp)
        memo[w_in & 1] = hit
    return hit[1]


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, act: int = ACT_RELU, kernel: Optional[str] = None
            ) -> torch.Tensor:
    """x: NHWC [n, h, w, cin]; w: [cout, 3, 3, cin] (same dtype); -> NHWC [n, h-2, w-2, cout].
    conv2's shape (bf16, 32 -> 64, an even number of input pixels) runs on the pixel-pair kernel unless
    ISTVT_CONV2_KERNEL / `kernel` names another construction."""
    dev = _chk(x, w, bias)
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    if tuple(w.shape) != (cout, 3, 3, cin) or w.dtype != x.dtype:
        raise ValueError("conv3x3: weight must be [cout, 3, 3, cin] in the activation dtype")
    y = torch.empty(n, h - 2, wd - 2, cout, dtype=x.dtype, device=dev)
    pair = _CONV2_PAIR if kernel is None else kernel == "pair"
    if (pair and x.dtype == torch.bfloat16 and cin == 32 and cout == 64 and (n * h * wd) % 2 == 0
            and act in (ACT_NONE, ACT_RELU) and x.is_contiguous() and w.is_contiguous()):
        wp = conv3x3_pair_weights(w, wd)
        with _launch(dev, "conv3x3", 2.0 * y.numel() * 9 * cin, _nbytes(x, w, y)):
            _lib.check(_lib.lib().istvt_conv3x3_pair_fwd(_ptr(x), _ptr(wp), _ptr(bias), _ptr(y), n, h, wd, act,
                                                         _stream(dev)), "istvt_conv3x3_pair_fwd")
        return y
    with _launch(dev, "conv3x3", 2.0 * y.numel() * 9 * cin, _nbytes(x, w, y)):
        _lib.check(_lib.lib().istvt_conv3x3_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), _dt(x), n, h, wd, cin, cout,
                                                act, _stream(dev)), "istvt_conv3x3_fwd")
    return y


def sepconv_fused_supported(c: int, n_out: int, w: int) -> bool:
    """Whether istvt_sepconv_fused_fwd takes this SeparableConv2d (same arithmetic as its host planner): 64-channel
    groups, <= 256 output channels, and the pointwise weights + two input ring stages within 227 KB of shared memory."""
    if c % 64 or c > 256 or n_out % 64 or not 64 <= n_out <= 256:
        return False
    strips = (w + 37) // 38
    pairs = ((w + strips - 1) // strips + 1) // 2
    stage = 5 * (2 * pairs + 2) * 128
    fixed = 2 * 16384 + 2 * 16384 + 4096 + 8192 + (c // 64) * n_out * 128 + 1024 + 256
    return (227 * 1024 - fixed) // stage >= 2


def sepconv_fused(x: torch.Tensor, dw: torch.Tensor, pw: torch.Tensor, bias: torch.Tensor, relu_in: bool,
                  act: int = ACT_NONE) -> torch.Tensor:
    """act(pointwise(depthwise3x3(relu_in ? relu(x) : x)) + bias) in one kernel (istvt_sepconv_fused_fwd).
    x: bf16 NHWC [n, h, w, c]; dw: fp32 [3, 3, c]; pw: bf16 [n_out, c]; -> bf16 NHWC [n, h, w, n_out]."""
    dev = _chk(x, dw, bias)
    _chk_rows(pw)
    n, h, w, c = x.shape
    n_out = pw.shape[0]
    if x.dtype != torch.bfloat16 or pw.dtype != torch.bfloat16 or pw.shape[1] != c or dw.numel() != 9 * c:
        raise ValueError("sepconv_fused: bf16 x [n, h, w, c], fp32 dw [3, 3, c], bf16 pw [n_out, c]")
    y = torch.empty(n, h, w, n_out, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "sepconv_fused", 2.0 * n * h * w * c * (9 + n_out), _nbytes(x, y, pw)):
        _lib.check(_lib.lib().istvt_sepconv_fused_fwd(_ptr(x), _ptr(dw), _ptr(pw), _ld(pw), _ptr(bias), _ptr(y), n, h, w, c,
                                                      n_out, int(relu_in), act, _stream(dev)), "istvt_sepconv_fused_fwd")
    return y


def conv_stem(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """x: fp32 NCHW [n, 3, h, w]; w: fp32 [32, 3, 3, 3]; -> NHWC [n, ho, wo, 32]."""
    dev = _chk(x, w, bias)
    if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("conv_stem expects an fp32 NCHW [n, 3, h, w] input")
    n, _, h, wd = x.shape
    cout = w.shape[0]
    ho, wo = (h - 3) // 2 + 1, (wd - 3) // 2 + 1
    y = torch.empty(n, ho, wo, cout, dtype=out_dtype, device=dev)
    with _launch(dev, "conv_stem", 2.0 * y.numel() * 27, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_conv_stem_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), _dt(y), n, h, wd, cout,
                                                  _stream(dev)), "istvt_conv_stem_fwd")
    return y


def conv_stem_u8(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """x: uint8 NHWC [n, h, w, 3] (decoded frames); w: fp32 [32, 3, 3, 3] with BN scale AND the input normalisation
    folded in (engine.fold_input_norm); -> NHWC [n, ho, wo, 32]."""
    dev = _chk(x, w, bias)
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
        raise ValueError("conv_stem_u8 expects a uint8 NHWC [n, h, w, 3] input")
    n, h, wd, _ = x.shape
    cout = w.shape[0]
    ho, wo = (h - 3) // 2 + 1, (wd - 3) // 2 + 1
    y = torch.empty(n, ho, wo, cout, dtype=out_dtype, device=dev)
    with _launch(dev, "conv_stem", 2.0 * y.numel() * 27, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_conv_stem_u8_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), _dt(y), n, h, wd, cout,
                                                     _stream(dev)), "istvt_conv_stem_u8_fwd")
    return y


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b (same shape / dtype, numel % 8 == 0): identity-skip residual of the Xception middle-flow blocks."""
    dev = _chk(a, b)
    if a.shape != b.shape or a.dtype != b.dtype or a.numel() % 8:
        raise ValueError("add: operands must have the same shape and dtype, numel % 8 == 0")
    y = torch.empty_like(a)
    with _launch(dev, "add", 0.0, _nbytes(a, b, y)):
        _lib.check(_lib.lib().istvt_add_fwd(_ptr(a), _ptr(b), _ptr(y), _dt(a), a.numel(), _stream(dev)), "istvt_add_fwd")
    return y


def pool_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True) -> torch.Tensor:
    """x: NHWC [n, h, w, c] -> fp32 [n, ncls] = Linear(mean_hw(relu(x))) (Xception.logits, xception.py:208-221)."""
    dev = _chk(x, w, bias)
    n, h, wd, c = x.shape
    if w.dtype != torch.float32 or w.dim() != 2 or w.shape[1] != c:
        raise ValueError("pool_linear: weight must be fp32 [ncls, c]")
    out = torch.empty(n, w.shape[0], dtype=torch.float32, device=dev)
    with _launch(dev, "pool_linear", 2.0 * n * c * w.shape[0], _nbytes(x)):
        _lib.check(_lib.lib().istvt_pool_linear_fwd(_ptr(x), _dt(x), _ptr(w), _ptr(bias), _ptr(out), n, h * wd, c,
                                                    w.shape[0], int(relu), _stream(dev)), "istvt_pool_linear_fwd")
    return out


def dwconv3x3(x: torch.Tensor, w: torch.Tensor, relu_in: bool) -> torch.Tensor:
    """x: NHWC; w: fp32 [3, 3, c]."""
    dev = _chk(x, w)
    n, h, wd, c = x.shape
    if tuple(w.shape) != (3, 3, c) or w.dtype != torch.float32:
        raise ValueError("dwconv3x3: weight must be fp32 [3, 3, c]")
    y = torch.empty_like(x)
    with _launch(dev, "dwconv3x3", 2.0 * y.numel() * 9, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_dwconv3x3_fwd(_ptr(x), _ptr(w), _ptr(y), _dt(x), n, h, wd, c, int(relu_in),
                                                  _stream(dev)), "istvt_dwconv3x3_fwd")
    return y


def subsample2(x: torch.Tensor) -> torch.Tensor:
    dev = _chk(x)
    n, h, wd, c = x.shape
    y = torch.empty(n, (h - 1) // 2 + 1, (wd - 1) // 2 + 1, c, dtype=x.dtype, device=dev)
    with _launch(dev, "subsample2", 0.0, 2 * _nbytes(y)):
        _lib.check(_lib.lib().istvt_subsample2_fwd(_ptr(x), _ptr(y), _dt(x), n, h, wd, c, _stream(dev)),
                   "istvt_subsample2_fwd")
    return y


def pool_add(x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    dev = _chk(x, skip)
    n, h, wd, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    if skip.numel() != n * ho * wo * c or skip.dtype != x.dtype:
        raise ValueError("pool_add: skip must be [n, ho, wo, c] in the activation dtype")
    y = torch.empty(n, ho, wo, c, dtype=x.dtype, device=dev)
    with _launch(dev, "pool_add", 0.0, _nbytes(x, skip, y)):
        _lib.check(_lib.lib().istvt_pool_add_fwd(_ptr(x), _ptr(skip), _ptr(y), _dt(x), n, h, wd, c, _stream(dev)),
                   "istvt_pool_add_fwd")
    return y


def pool_add_tokens(x: torch.Tensor, skip: torch.Tensor, pos_emb: torch.Tensor, tokens: torch.Tensor,
                    batch: int, t: int) -> None:
    """Writes tokens[b, f+1, 1+p, :] = maxpool(x)+skip+pos_emb[f, 1+p, :]; tokens: fp32 [B, T+1, P, C]."""
    dev = _chk(x, skip, pos_emb, tokens)
    n, h, wd, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    if n != batch * t or tokens.dtype != torch.float32 or pos_emb.dtype != torch.float32:
        raise ValueError("pool_add_tokens: bad batch/frames or dtypes")
    if tokens.numel() != batch * (t + 1) * (ho * wo + 1) * c or pos_emb.numel() != t * (ho * wo + 1) * c:
        raise ValueError("pool_add_tokens: token / pos_emb buffers have the wrong size")
    with _launch(dev, "pool_add", 0.0, _nbytes(x, skip) + 2 * skip.numel() * 4):
        _lib.check(_lib.lib().istvt_pool_add_tokens_fwd(_ptr(x), _ptr(skip), _ptr(pos_emb), _ptr(tokens), _dt(x),
                                                        batch, t, h, wd, c, _stream(dev)),
                   "istvt_pool_add_tokens_fwd")


def token_fill(tokens: torch.Tensor, space_token: torch.Tensor, temporal_token: torch.Tensor,
               pos_emb: torch.Tensor) -> None:
    dev = _chk(tokens, space_token, temporal_token, pos_emb)
    b, f, p, d = tokens.shape
    with _launch(dev, "token_fill", 0.0, b * (p + f - 1) * d * 4):
        _lib.check(_lib.lib().istvt_token_fill_fwd(_ptr(tokens), _ptr(space_token), _ptr(temporal_token),
                                                   _ptr(pos_emb), b, f - 1, p, d, _stream(dev)),
                   "istvt_token_fill_fwd")


def attn_temporal(qk: torch.Tensor, v: torch.Tensor, batch: int, frames: int, tokens: int, heads: int,
                  scale: float, want_probs: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    dev = _chk(qk, v)
    inner = heads * 64
    rows = batch * frames * tokens
    if qk.numel() != rows * 2 * inner or v.numel() != rows * inner or qk.dtype != v.dtype:
        raise ValueError("attn_temporal: qk must be [rows, 2*heads*64] and v [rows, heads*64]")
    out = torch.empty(rows, inner, dtype=qk.dtype, device=dev)
    probs = torch.empty(batch, heads, tokens, frames, frames, dtype=torch.float32, device=dev) if want_probs else None
    with _launch(dev, "attn_temporal", 4.0 * batch * tokens * heads * frames * frames * 64, _nbytes(qk, v, out, probs)):
        _lib.check(_lib.lib().istvt_attn_temporal_fwd(_ptr(qk), _ptr(v), _ptr(out), _ptr(probs), _dt(qk), batch,
                                                      frames, tokens, heads, scale, _stream(dev)),
                   "istvt_attn_temporal_fwd")
    return out, probs


def attn_spatial(qkv: torch.Tensor, batch_frames: int, tokens: int, heads: int, scale: float,
                 want_probs: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    dev = _chk(qkv)
    inner = heads * 64
    rows = batch_frames * tokens
    if qkv.numel() != rows * 3 * inner:
        raise ValueError("attn_spatial: qkv must be [rows, 3*heads*64]")
    out = torch.empty(rows, inner, dtype=qkv.dtype, device=dev)
    probs = torch.empty(batch_frames, heads, tokens, tokens, dtype=torch.float32, device=dev) if want_probs else None
    with _launch(dev, "attn_spatial", 4.0 * batch_frames * heads * tokens * tokens * 64, _nbytes(qkv, out, probs)):
        _lib.check(_lib.lib().istvt_attn_spatial_fwd(_ptr(qkv), _ptr(out), _ptr(probs), _dt(qkv), batch_frames,
                                                     tokens, heads, scale, _stream(dev)), "istvt_attn_spatial_fwd")
    return out, probs


def attn_joint(qkv: torch.Tensor, batch: int, tokens: int, heads: int, scale: float) -> torch.Tensor:
    """Joint self-attention over all `tokens` of each of the `batch` sequences, any sequence length (module.py:53-63).
    qkv: [batch*tokens, 3*heads*64] -> [batch*tokens, heads*64]; keys are streamed with an online softmax."""
    dev = _chk(qkv)
    inner = heads * 64
    rows = batch * tokens
    if qkv.numel() != rows * 3 * inner:
        raise ValueError("attn_joint: qkv must be [batch*tokens, 3*heads*64]")
    out = torch.empty(rows, inner, dtype=qkv.dtype, device=dev)
    with _launch(dev, "attn_joint", 4.0 * batch * heads * tokens * tokens * 64, _nbytes(qkv, out)):
        _lib.check(_lib.lib().istvt_attn_joint_fwd(_ptr(qkv), _ptr(out), _dt(qkv), batch, tokens, heads, scale,
                                                   _stream(dev)), "istvt_attn_joint_fwd")
    return out


def token_build(src: torch.Tensor, cls: torch.Tensor, pos: Optional[torch.Tensor], sequences: int, n: int,
                pos_period: int = 1) -> torch.Tensor:
    """Class token + n patch rows (+ positional embedding) per sequence -> fp32 tokens [sequences, n+1, dim]
    (vivit.py:64-66, :73-74, :183-185).  src: [sequences*n, dim] bf16/fp32; cls fp32 [dim]; pos fp32
    [pos_period, n+1, dim] or None."""
    dev = _chk(src, cls, pos)
    dim = src.shape[-1]
    if src.numel() != sequences * n * dim or cls.numel() != dim or cls.dtype != torch.float32:
        raise ValueError("token_build: src must be [sequences*n, dim] and cls fp32 [dim]")
    if pos is not None and (pos.dtype != torch.float32 or pos.numel() != pos_period * (n + 1) * dim):
        raise ValueError("token_build: pos must be fp32 [pos_period, n+1, dim]")
    tokens = torch.empty(sequences, n + 1, dim, dtype=torch.float32, device=dev)
    with _launch(dev, "token_build", 0.0, _nbytes(src, tokens) + (0 if pos is None else tokens.numel() * 4)):
        _lib.check(_lib.lib().istvt_token_build_fwd(_ptr(src), _dt(src), _ptr(cls), _ptr(pos), _ptr(tokens), sequences,
                                                    n, dim, pos_period, _stream(dev)), "istvt_token_build_fwd")
    return tokens


def mean_rows(x: torch.Tensor, sequences: int, n: int) -> torch.Tensor:
    """x fp32 [sequences, n, dim] -> fp32 [sequences, dim], the mean over the n token rows (vivit.py:79)."""
    dev = _chk(x)
    dim = x.shape[-1]
    if x.dtype != torch.float32 or x.numel() != sequences * n * dim:
        raise ValueError("mean_rows: x must be fp32 [sequences, n, dim]")
    out = torch.empty(sequences, dim, dtype=torch.float32, device=dev)
    with _launch(dev, "mean_rows", 0.0, _nbytes(x, out)):
        _lib.check(_lib.lib().istvt_mean_rows_fwd(_ptr(x), _ptr(out), sequences, n, dim, _stream(dev)),
                   "istvt_mean_rows_fwd")
    return out


def head(tokens: torch.Tensor, norm_g, norm_b, head_g, head_b, head_w, head_bias, eps: float = 1e-5) -> torch.Tensor:
    """tokens: fp32 [B, F, P, D] -> logits fp32 [B, 1] from token (0, 0)."""
    dev = _chk(tokens, norm_g, norm_b, head_g, head_b, head_w, head_bias)
    b, f, p, d = tokens.shape
    logits = torch.empty(b, 1, dtype=torch.float32, device=dev)
    with _launch(dev, "head", 0.0, b * d * 4):
        _lib.check(_lib.lib().istvt_head_fwd(_ptr(tokens), f * p, _ptr(norm_g), _ptr(norm_b), _ptr(head_g),
                                             _ptr(head_b), _ptr(head_w), _ptr(head_bias), _ptr(logits), b, d, eps,
                                             _stream(dev)), "istvt_head_fwd")
    return logits


# ----------------------------------------------------------------------------------------------
# training step (backward kernels, optimizer)
# ----------------------------------------------------------------------------------------------
def attn_spatial_lse(qkv: torch.Tensor, batch_frames: int, tokens: int, heads: int, scale: float
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Training-mode spatial attention: (out [rows, heads*64] bf16, lse [batch_frames, heads, tokens] fp32)."""
    dev = _chk(qkv)
    inner = heads * 64
    rows = batch_frames * tokens
    if qkv.dtype != torch.bfloat16 or qkv.numel() != rows * 3 * inner:
        raise ValueError("attn_spatial_lse: qkv must be bf16 [rows, 3*heads*64]")
    out = torch.empty(rows, inner, dtype=qkv.dtype, device=dev)
    lse = torch.empty(batch_frames, heads, tokens, dtype=torch.float32, device=dev)
    with _launch(dev, "attn_spatial", 4.0 * batch_frames * heads * tokens * tokens * 64, _nbytes(qkv, out, lse)):
        _lib.check(_lib.lib().istvt_attn_spatial_fwd_lse(_ptr(qkv), _ptr(out), _ptr(lse), batch_frames, tokens, heads,
                                                         scale, _stream(dev)), "istvt_attn_spatial_fwd_lse")
    return out, lse


def attn_spatial_bwd(qkv: torch.Tensor, o: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, batch_frames: int,
                     tokens: int, heads: int, scale: float, scratch: Optional[torch.Tensor] = None,
                     cam: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cam (optional, fp32 [batch_frames, tokens, tokens], zeroed by the caller) += relu(dA o A) / heads."""
    dev = _chk(qkv, o, dout, lse, scratch, cam)
    inner = heads * 64
    rows = batch_frames * tokens
    if any(t.dtype != torch.bfloat16 for t in (qkv, o, dout)) or lse.dtype != torch.float32:
        raise ValueError("attn_spatial_bwd: qkv / o / dout must be bf16 and lse fp32")
    if qkv.numel() != rows * 3 * inner or o.numel() != rows * inner or dout.numel() != rows * inner:
        raise ValueError("attn_spatial_bwd: shape mismatch")
    dqkv = torch.empty_like(qkv)
    if scratch is None:
        scratch = torch.empty(rows, inner, dtype=torch.float32, device=dev)
    with _launch(dev, "attn_spatial_bwd", 10.0 * batch_frames * heads * tokens * tokens * 64,
                 _nbytes(qkv, o, dout, dqkv)):
        if cam is None:
            _lib.check(_lib.lib().istvt_attn_spatial_bwd(_ptr(qkv), _ptr(o), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                                         _ptr(scratch), batch_frames, tokens, heads, scale,
                                                         _stream(dev)), "istvt_attn_spatial_bwd")
        else:
            if cam.dtype != torch.float32 or cam.numel() != batch_frames * tokens * tokens:
                raise ValueError("attn_spatial_bwd: cam must be fp32 [batch_frames, tokens, tokens]")
            _lib.check(_lib.lib().istvt_attn_spatial_bwd_cam(_ptr(qkv), _ptr(o), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                                             _ptr(scratch), _ptr(cam), batch_frames, tokens, heads,
                                                             scale, _stream(dev)), "istvt_attn_spatial_bwd_cam")
    return dqkv


def attn_temporal_bwd(qk: torch.Tensor, v: torch.Tensor, dout: torch.Tensor, batch: int, frames: int, tokens: int,
                      heads: int, scale: float, cam: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """cam (optional, fp32 [batch, tokens, frames, frames], zeroed by the caller) += relu(dA o A) / heads."""
    dev = _chk(qk, v, dout, cam)
    if any(t.dtype != torch.bfloat16 for t in (qk, v, dout)):
        raise ValueError("attn_temporal_bwd: bf16 only")
    dqk, dv = torch.empty_like(qk), torch.empty_like(v)
    with _launch(dev, "attn_temporal_bwd", 10.0 * batch * tokens * heads * frames * frames * 64,
                 _nbytes(qk, v, dout, dqk, dv)):
        if cam is None:
            _lib.check(_lib.lib().istvt_attn_temporal_bwd(_ptr(qk), _ptr(v), _ptr(dout), _ptr(dqk), _ptr(dv), batch,
                                                          frames, tokens, heads, scale, _stream(dev)),
                       "istvt_attn_temporal_bwd")
        else:
            if cam.dtype != torch.float32 or cam.numel() != batch * tokens * frames * frames:
                raise ValueError("attn_temporal_bwd: cam must be fp32 [batch, tokens, frames, frames]")
            _lib.check(_lib.lib().istvt_attn_temporal_bwd_cam(_ptr(qk), _ptr(v), _ptr(dout), _ptr(dqk), _ptr(dv), _ptr(cam),
                                                              batch, frames, tokens, heads, scale, _stream(dev)),
                       "istvt_attn_temporal_bwd_cam")
    return dqk, dv


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor,
                  g_accum: Optional[torch.Tensor] = None, g_bf16: Optional[torch.Tensor] = None,
                  dy2: Optional[torch.Tensor] = None, frames: int = 0, tokens_per_frame: int = 0,
                  eps: float = 1e-5, out_colsum: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Returns dx (bf16) unless `g_accum` (fp32, += dx) is given.  `out_colsum` (fp32 [dim], +=): column sums of the
    rows produced (updated g_accum, or dx) — the bias gradient of the Linear whose output gradient they are."""
    dev = _chk(gamma, dgamma, dbeta, out_colsum)
    _chk_rows(dy, x, g_accum, g_bf16, dy2)
    dim = x.shape[-1]
    rows = x.numel() // dim
    if dy.dtype != torch.bfloat16 or dy.numel() != x.numel():
        raise ValueError("layernorm_bwd: dy must be bf16 with x's shape")
    if dy2 is not None and _ld(dy2) != _ld(dy):
        raise ValueError("layernorm_bwd: dy and dy2 must have the same row pitch")
    # dx is the A operand of the next GEMM / weight-gradient GEMM: 128-byte aligned row pitch when dy has one
    dx = None
    if g_accum is None:
        dx = empty_rows((rows, dim), torch.bfloat16, dev) if _ld(dy) != dim else torch.empty(rows, dim, dtype=torch.bfloat16,
                                                                                             device=dev)
    with _launch(dev, "layernorm_bwd", 0.0, _nbytes(dy, x, dx, dy2, g_bf16) + 2 * _nbytes(g_accum)):
        _lib.check(_lib.lib().istvt_layernorm_bwd_ld(
            _ptr(dy), _ptr(dy2), _ld(dy), frames, tokens_per_frame, _ptr(x), _dt(x), _ld(x), _ptr(gamma), _ptr(g_accum),
            _ld(g_accum) if g_accum is not None else dim, _ptr(g_bf16), _ld(g_bf16) if g_bf16 is not None else dim,
            _ptr(dx), _ld(dx) if dx is not None else dim, _ptr(dgamma), _ptr(dbeta), _ptr(out_colsum), rows, dim, eps,
            _stream(dev)),
            "istvt_layernorm_bwd_ld")
    return dx


def gelu(x: torch.Tensor) -> torch.Tensor:
    dev = _chk(x)
    y = torch.empty_like(x)
    with _launch(dev, "gelu", 0.0, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_gelu_fwd(_ptr(x), _ptr(y), x.numel(), _stream(dev)), "istvt_gelu_fwd")
    return y


def gelu_bwd(dy: torch.Tensor, x: torch.Tensor, colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = dy * gelu'(x); `colsum` (fp32 [cols], +=): column sums of dx (x: [rows, cols]) from the same pass."""
    dev = _chk(dy, x, colsum)
    dx = torch.empty_like(x)
    with _launch(dev, "gelu_bwd", 0.0, _nbytes(dy, x, dx)):
        if colsum is None:
            _lib.check(_lib.lib().istvt_gelu_bwd(_ptr(dy), _ptr(x), _ptr(dx), x.numel(), _stream(dev)), "istvt_gelu_bwd")
        else:
            cols = x.shape[-1]
            _lib.check(_lib.lib().istvt_gelu_bwd_colsum(_ptr(dy), _ptr(x), _ptr(dx), _ptr(colsum), x.numel() // cols, cols,
                                                        _stream(dev)), "istvt_gelu_bwd_colsum")
    return dx


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16; `out` (or x) may be a row-pitched matrix (ops.empty_rows)."""
    dev = _chk_rows(x, out)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "cast", 0.0, _nbytes(x, out)):
        if x.is_contiguous() and out.is_contiguous():
            _lib.check(_lib.lib().istvt_cast_f32_bf16(_ptr(x), _ptr(out), x.numel(), _stream(dev)), "istvt_cast_f32_bf16")
        else:
            cols = x.shape[-1]
            _lib.check(_lib.lib().istvt_cast_f32_bf16_rows(_ptr(x), _ld(x), _ptr(out), _ld(out), x.numel() // cols, cols,
                                                           _stream(dev)), "istvt_cast_f32_bf16_rows")
    return out


def transpose(x: torch.Tensor, colsum: Optional[torch.Tensor] = None, aligned: bool = False) -> torch.Tensor:
    """x: bf16 [M, C] -> bf16 [C, ld] with ld = M rounded up to 8 (pad zero); colsum (fp32 [C]) += column sums.
    aligned: ld = the 128-byte aligned row pitch of M instead (the result is a GEMM operand with K = M)."""
    dev = _chk(x, colsum)
    m, c = x.shape
    ld = max(row_pitch(m), (m + 7) // 8 * 8) if aligned else (m + 7) // 8 * 8
    out = torch.empty(c, ld, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "transpose", 0.0, 2 * _nbytes(x)):
        _lib.check(_lib.lib().istvt_transpose_colsum(_ptr(x), _ptr(out), _ptr(colsum), m, c, ld, _stream(dev)),
                   "istvt_transpose_colsum")
    return out[:, :m] if aligned and ld != m else out


def gemm_wgrad(dyt: torch.Tensor, xt: torch.Tensor, rows: int, dw: torch.Tensor) -> None:
    """dw[N, K] (fp32) += dyt[N, :rows] . xt[K, :rows]^T  (both operands K-major over the token rows)."""
    dev = _chk(dyt, xt, dw)
    n, ld = dyt.shape
    k = xt.shape[0]
    if xt.shape[1] != ld or dw.dtype != torch.float32 or dw.numel() != n * k:
        raise ValueError("gemm_wgrad: operand shapes do not match")
    with _launch(dev, "gemm_wgrad", 2.0 * n * k * rows, _nbytes(dyt, xt) + 2 * _nbytes(dw)):
        _lib.check(_lib.lib().istvt_gemm_splitk_accum(_ptr(dyt), ld, _ptr(xt), ld, _ptr(dw), k, n, k, rows,
                                                      _stream(dev)), "istvt_gemm_splitk_accum")


def head_bwd(tokens: torch.Tensor, dlogits: torch.Tensor, norm_g, norm_b, head_g, head_b, head_w, g: torch.Tensor,
             d_norm_g, d_norm_b, d_head_g, d_head_b, d_head_w, d_head_bias, eps: float = 1e-5) -> None:
    dev = _chk(tokens, dlogits, g)
    b, f, p, d = tokens.shape
    with _launch(dev, "head_bwd", 0.0, 0.0):
        _lib.check(_lib.lib().istvt_head_bwd(_ptr(tokens), f * p, _ptr(dlogits), _ptr(norm_g), _ptr(norm_b), _ptr(head_g),
                                             _ptr(head_b), _ptr(head_w), _ptr(g), _ptr(d_norm_g), _ptr(d_norm_b),
                                             _ptr(d_head_g), _ptr(d_head_b), _ptr(d_head_w), _ptr(d_head_bias), b, d,
                                             eps, _stream(dev)), "istvt_head_bwd")


def token_bwd(g: torch.Tensor, d_pos: torch.Tensor, d_space: torch.Tensor, d_temporal: torch.Tensor) -> None:
    dev = _chk(g, d_pos, d_space, d_temporal)
    b, f, p, d = g.shape
    with _launch(dev, "token_bwd", 0.0, _nbytes(g)):
        _lib.check(_lib.lib().istvt_token_bwd(_ptr(g), _ptr(d_pos), _ptr(d_space), _ptr(d_temporal), b, f - 1, p, d,
                                              _stream(dev)), "istvt_token_bwd")


def adamw_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
               betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, step: int = 1,
               grad_scale: float = 1.0) -> None:
    dev = _chk(params, grads, exp_avg, exp_avg_sq)
    n = params.numel()
    with _launch(dev, "adamw", 0.0, 7 * n * 4):
        _lib.check(_lib.lib().istvt_adamw_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), n, lr,
                                               betas[0], betas[1], eps, weight_decay, step, grad_scale, _stream(dev)),
                   "istvt_adamw_step")


# ---- entry flow, training mode ----
def conv_stem_raw(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """conv1 without BN / ReLU: fp32 NCHW [n, 3, h, w] -> bf16 NHWC [n, ho, wo, 32]."""
    dev = _chk(x, w)
    n, _, h, wd = x.shape
    ho, wo = (h - 3) // 2 + 1, (wd - 3) // 2 + 1
    y = torch.empty(n, ho, wo, w.shape[0], dtype=torch.bfloat16, device=dev)
    zero = torch.zeros(w.shape[0], dtype=torch.float32, device=dev)
    with _launch(dev, "conv_stem", 2.0 * y.numel() * 27, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_conv_stem_raw_fwd(_ptr(x), _ptr(w), _ptr(zero), _ptr(y), BF16, n, h, wd, w.shape[0],
                                                      _stream(dev)), "istvt_conv_stem_raw_fwd")
    return y


class BNState:
    """Per-layer BatchNorm batch statistics kept for the backward."""
    __slots__ = ("scale", "shift", "mean", "rstd", "m", "c")


def batchnorm_train(x: torch.Tensor, bn, relu: bool, update_running: bool = True) -> Tuple[torch.Tensor, BNState]:
    """x: bf16 [..., c] raw convolution output; bn: nn.BatchNorm2d (parameters fp32).  y = [relu](bn_batch(x))."""
    dev = _chk(x)
    c = x.shape[-1]
    m = x.numel() // c
    stats = torch.zeros(2, c, dtype=torch.float32, device=dev)
    st = BNState()
    st.m, st.c = m, c
    buf = torch.empty(4, c, dtype=torch.float32, device=dev)
    st.scale, st.shift, st.mean, st.rstd = buf[0], buf[1], buf[2], buf[3]
    s = _stream(dev)
    with _launch(dev, "bn_stats", 0.0, _nbytes(x)):
        _lib.check(_lib.lib().istvt_bn_stats_fwd(_ptr(x), _ptr(stats[0]), _ptr(stats[1]), m, c, s), "istvt_bn_stats_fwd")
    rm = bn.running_mean if update_running else None
    rv = bn.running_var if update_running else None
    _lib.check(_lib.lib().istvt_bn_finalize_fwd(_ptr(stats[0]), _ptr(stats[1]), _ptr(bn.weight.data), _ptr(bn.bias.data),
                                                _ptr(st.scale), _ptr(st.shift), _ptr(st.mean), _ptr(st.rstd),
                                                _ptr(rm), _ptr(rv), m, c, bn.eps, bn.momentum, s),
               "istvt_bn_finalize_fwd")
    if update_running:
        bn.num_batches_tracked += 1
    y = torch.empty_like(x)
    with _launch(dev, "bn_apply", 0.0, _nbytes(x, y)):
        _lib.check(_lib.lib().istvt_bn_apply_fwd(_ptr(x), _ptr(st.scale), _ptr(st.shift), _ptr(y), m, c, int(relu), s),
                   "istvt_bn_apply_fwd")
    return y, st


def batchnorm_bwd(dy: torch.Tensor, x: torch.Tensor, st: BNState, dgamma: torch.Tensor, dbeta: torch.Tensor,
                  relu: bool) -> torch.Tensor:
    """dgamma / dbeta are ACCUMULATED into (gradient accumulation over micro-batches keeps working): the kernel reduces
    this layer's sums into a zeroed per-call scratch — it derives dx from exactly those sums — and the scratch is then
    added to the caller's slots."""
    dev = _chk(dy, x, dgamma, dbeta)
    dx = torch.empty_like(x)
    sums = torch.zeros(2, st.c, dtype=torch.float32, device=dev)
    with _launch(dev, "bn_bwd", 0.0, 2 * _nbytes(dy, x) + _nbytes(dx)):
        _lib.check(_lib.lib().istvt_bn_bwd(_ptr(dy), _ptr(x), _ptr(st.scale), _ptr(st.shift), _ptr(st.mean), _ptr(st.rstd),
                                           _ptr(sums[0]), _ptr(sums[1]), _ptr(dx), st.m, st.c, int(relu), _stream(dev)),
                   "istvt_bn_bwd")
    dgamma.add_(sums[0])
    dbeta.add_(sums[1])
    return dx


def pool_add_idx(x: torch.Tensor, skip: torch.Tensor, tokens: Optional[torch.Tensor] = None,
                 pos_emb: Optional[torch.Tensor] = None, t_frames: int = 1):
    """Training-mode pool+add: returns (y or None, argmax uint8 [n, ho, wo, c])."""
    dev = _chk(x, skip, tokens, pos_emb)
    n, h, wd, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    amax = torch.empty(n, ho, wo, c, dtype=torch.uint8, device=dev)
    y = None if tokens is not None else torch.empty(n, ho, wo, c, dtype=x.dtype, device=dev)
    with _launch(dev, "pool_add", 0.0, _nbytes(x, skip, y, amax)):
        _lib.check(_lib.lib().istvt_pool_add_idx_fwd(_ptr(x), _ptr(skip), _ptr(y), _ptr(pos_emb), _ptr(tokens), _ptr(amax),
                                                     n, t_frames, h, wd, c, _stream(dev)), "istvt_pool_add_idx_fwd")
    return y, amax


def pool_bwd(dy: torch.Tensor, amax: torch.Tensor, h: int, wd: int) -> torch.Tensor:
    dev = _chk(dy, amax)
    n, _, _, c = dy.shape
    dx = torch.empty(n, h, wd, c, dtype=dy.dtype, device=dev)
    with _launch(dev, "pool_bwd", 0.0, _nbytes(dy, amax, dx)):
        _lib.check(_lib.lib().istvt_pool_bwd(_ptr(dy), _ptr(amax), _ptr(dx), n, h, wd, c, _stream(dev)), "istvt_pool_bwd")
    return dx


def token_grad_gather(g: torch.Tensor) -> torch.Tensor:
    """g: fp32 [B, T+1, P, C] -> bf16 NHWC [B*T, side, side, C] (gradient of block 3's output)."""
    dev = _chk(g)
    b, f, p, c = g.shape
    side = int(round((p - 1) ** 0.5))
    out = torch.empty(b * (f - 1), side, side, c, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "token_grad_gather", 0.0, _nbytes(out) * 3):
        _lib.check(_lib.lib().istvt_token_grad_gather(_ptr(g), _ptr(out), b, f - 1, p, c, _stream(dev)),
                   "istvt_token_grad_gather")
    return out


def dwconv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, relu_in: bool) -> None:
    """dw: fp32 [3, 3, c] (+=)."""
    dev = _chk(x, dy, dw)
    n, h, wd, c = x.shape
    with _launch(dev, "dwconv_wgrad", 2.0 * x.numel() * 9, _nbytes(x, dy)):
        _lib.check(_lib.lib().istvt_dwconv3x3_wgrad(_ptr(x), _ptr(dy), _ptr(dw), n, h, wd, c, int(relu_in), _stream(dev)),
                   "istvt_dwconv3x3_wgrad")


def block_input_grad(d_main: torch.Tensor, x_in: Optional[torch.Tensor], d_skip: torch.Tensor, relu_in: bool
                     ) -> torch.Tensor:
    dev = _chk(d_main, x_in, d_skip)
    n, h, wd, c = d_main.shape
    dx = torch.empty_like(d_main)
    with _launch(dev, "block_input_grad", 0.0, _nbytes(d_main, x_in if relu_in else None, d_skip, dx)):
        _lib.check(_lib.lib().istvt_block_input_grad(_ptr(d_main), _ptr(x_in), _ptr(d_skip), _ptr(dx), n, h, wd, c,
                                                     int(relu_in), _stream(dev)), "istvt_block_input_grad")
    return dx


def im2col_t(x: torch.Tensor) -> torch.Tensor:
    """NHWC bf16 [n, h, w, cin] -> [9*cin, ld] (3x3 valid taps, K-major over output pixels)."""
    dev = _chk(x)
    n, h, wd, cin = x.shape
    m = n * (h - 2) * (wd - 2)
    ld = (m + 7) // 8 * 8
    out = torch.empty(9 * cin, ld, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "im2col_t", 0.0, _nbytes(out) * 2):
        _lib.check(_lib.lib().istvt_im2col_t(_ptr(x), _ptr(out), n, h, wd, cin, ld, _stream(dev)), "istvt_im2col_t")
    return out


def im2col_t_stem(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW [n, 3, h, w] -> bf16 [32, ld] (27 stride-2 taps + 5 zero rows)."""
    dev = _chk(x)
    n, _, h, wd = x.shape
    m = n * ((h - 3) // 2 + 1) * ((wd - 3) // 2 + 1)
    ld = (m + 7) // 8 * 8
    out = torch.empty(32, ld, dtype=torch.bfloat16, device=dev)
    with _launch(dev, "im2col_t", 0.0, _nbytes(x, out)):
        _lib.check(_lib.lib().istvt_im2col_t_stem(_ptr(x), _ptr(out), n, h, wd, ld, _stream(dev)), "istvt_im2col_t_stem")
    return out


def rollout_row(v: torch.Tensor, cmat: torch.Tensor) -> None:
    """v: fp32 [n, L] (in place) <- v (I + cmat), cmat fp32 [n, L, L]."""
    dev = _chk(v, cmat)
    n, ln = v.shape
    if cmat.numel() != n * ln * ln or v.dtype != torch.float32 or cmat.dtype != torch.float32:
        raise ValueError("rollout_row: cmat must be fp32 [n, L, L]")
    with _launch(dev, "rollout_row", 2.0 * n * ln * ln, _nbytes(cmat)):
        _lib.check(_lib.lib().istvt_rollout_row(_ptr(v), _ptr(cmat), n, ln, _stream(dev)), "istvt_rollout_row")


def gather_rows(src: torch.Tensor, n_outer: int, outer_stride: int, rows: int, row_stride: int, width: int
                ) -> torch.Tensor:
    """dst[o, r, :width] = src.flatten()[o*outer_stride + r*row_stride : ... + width] (element units) -> [n_outer*rows, width]."""
    dev = _chk(src)
    es = src.element_size()
    dst = torch.empty(n_outer * rows, width, dtype=src.dtype, device=dev)
    with _launch(dev, "gather_rows", 0.0, 2 * _nbytes(dst)):
        _lib.check(_lib.lib().istvt_gather_rows(_ptr(src), _ptr(dst), n_outer, outer_stride * es, rows, row_stride * es,
                                                width * es, _stream(dev)), "istvt_gather_rows")
    return dst


def colsum(x: torch.Tensor, out: torch.Tensor) -> None:
    """out (fp32 [C]) += column sums of x (bf16 [M, C])."""
    dev = _chk(out)
    _chk_rows(x)
    m, c = x.shape
    with _launch(dev, "colsum", 0.0, _nbytes(x)):
        _lib.check(_lib.lib().istvt_colsum_ld(_ptr(x), _ld(x), _ptr(out), m, c, _stream(dev)), "istvt_colsum_ld")


def wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, bias_grad: Optional[torch.Tensor] = None) -> None:
    """dw[N, K] (fp32) += dy[rows, N]^T x[rows, K]; bias_grad (fp32 [N]) += column sums of dy.
    Operands are read in place (MN-major tcgen05 tiles); tiny problems (rows <= 64) take the transposed-copy path."""
    dev = _chk(dw, bias_grad)
    _chk_rows(dy, x)
    rows, n = dy.shape
    k = x.shape[1]
    if x.shape[0] != rows or dw.numel() != n * k or dw.dtype != torch.float32:
        raise ValueError("wgrad: operand shapes do not match")
    if rows <= 64:
        gemm_wgrad(transpose(dy.contiguous(), colsum=bias_grad), transpose(x.contiguous()), rows, dw)
        return
    if bias_grad is not None:
        colsum(dy, bias_grad)
    with _launch(dev, "gemm_wgrad", 2.0 * n * k * rows, _nbytes(dy, x) + 2 * _nbytes(dw)):
        _lib.check(_lib.lib().istvt_gemm_wgrad_accum(_ptr(dy), _ld(dy), _ptr(x), _ld(x), _ptr(dw), k, rows, n, k,
                                                     _stream(dev)), "istvt_gemm_wgrad_accum")
