"""conv2 on pixel pairs (csrc/gemm_tcgen05.cu: conv3x3_pair_launch / conv3x3_pair_pack_kernel; replaces conv2 + bn2 + relu,
/root/reference/network/xception.py:122-123,198-200): the formulation itself, restated in torch (tests/torch_ops.py), must
equal a plain 3x3 convolution for both parities of the input width.  The GPU check `conv3x3` compares the CUDA kernels
with the same convolution and the CUDA weight packing with this mirror bit for bit."""
import torch
import torch.nn.functional as F

import torch_ops


def _case(n, h, w, act, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, 32, generator=g)
    wt = torch.randn(64, 3, 3, 32, generator=g) / 17.0
    b = torch.randn(64, generator=g) * 0.1
    ref = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(0, 3, 1, 2), b)
    ref = (ref.relu() if act else ref).permute(0, 2, 3, 1)
    got = torch_ops.conv3x3_pair_ref(x, wt, b, act)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())


def test_pair_formulation_odd_width():
    _case(2, 9, 11, 1, 0)          # W odd: 7 k-blocks, the ky = 1 row straddles three pairs
    _case(2, 3, 3, 0, 1)           # 1 x 1 outputs


def test_pair_formulation_even_width():
    _case(3, 8, 12, 1, 2)          # W even: 6 k-blocks
    _case(1, 6, 4, 0, 3)


def test_pair_tap_table():
    assert [t[3] for t in torch_ops.conv3x3_pair_taps(149)] == [0, 1, 74, 75, 76, 149, 150]
    assert [t[3] for t in torch_ops.conv3x3_pair_taps(12)] == [0, 1, 6, 7, 12, 13]
    wp = torch_ops.conv3x3_pair_pack_ref(torch.ones(64, 3, 3, 32), 149)
    assert wp.shape == (128, 448)
    # every output pixel sees each of its 9 taps exactly once per input channel
    assert int(wp[:64].sum()) == 64 * 9 * 32 and int(wp[64:].sum()) == 64 * 9 * 32
