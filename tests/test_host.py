"""CPU: host-side logic — module tree / state_dict schema, registry, error behaviour, BN folding."""
import pytest
import torch
import torch.nn.functional as F

from helpers import pkg


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return pkg().XceptionVidTr().eval()


def test_state_dict_schema(model):
    sd = model.state_dict()
    assert len(sd) == 489
    assert sd["xcep.model.conv1.weight"].shape == (32, 3, 3, 3)
    assert sd["xcep.model.block1.rep.0.conv1.weight"].shape == (64, 1, 3, 3)
    assert sd["xcep.model.block1.rep.3.pointwise.weight"].shape == (128, 128, 1, 1)
    assert sd["xcep.model.block3.skip.weight"].shape == (728, 256, 1, 1)
    assert sd["xcep.model.block3.rep.4.pointwise.weight"].shape == (728, 728, 1, 1)
    assert sd["xcep.model.block12.rep.4.pointwise.weight"].shape == (1024, 728, 1, 1)
    assert sd["xcep.model.last_linear.1.weight"].shape == (2, 2048)
    assert sd["vit.pos_embedding"].shape == (1, 6, 362, 728)
    assert sd["vit.transformer.layers.11.0.fn.to_qk.weight"].shape == (1024, 728)
    assert sd["vit.transformer.layers.0.0.fn.to_v.weight"].shape == (512, 728)
    assert sd["vit.transformer.layers.0.1.fn.to_qkv.weight"].shape == (1536, 728)
    assert sd["vit.transformer.layers.0.2.fn.net.3.weight"].shape == (728, 2912)
    assert sd["vit.mlp_head.1.weight"].shape == (1, 728)
    assert sum(p.numel() for p in model.parameters()) == 109172051


def test_registry():
    m = pkg()
    assert isinstance(m.model_selection("resnet_3d", num_out_classes=1), m.XceptionVidTr)
    x = m.model_selection("xception", num_out_classes=2, dropout=0.5)
    assert isinstance(x, m.TransferModel) and hasattr(x.model, "block12")
    with pytest.raises(NotImplementedError):
        m.model_selection("jigsaw_multi_xcep_adv", num_out_classes=2)


def test_no_cpu_fallback(model):
    with pytest.raises(ValueError, match="CUDA"):
        model(torch.zeros(1, 6, 3, 300, 300))
    with pytest.raises(ValueError):
        model(torch.zeros(1, 6, 300, 300))


def test_long_clip_ctor():
    m = pkg().XceptionVidTr(num_frames=32)
    assert m.vit.pos_embedding.shape == (1, 32, 362, 728)


def test_bn_folding_matches_torch(model):
    """pack_entry's folded conv2 weights/bias reproduce conv2 -> bn2 (eval) on CPU."""
    eng = __import__("importlib").import_module("2023-tifs-istvt_b200.engine")
    x = model.xcep.model
    with torch.no_grad():
        x.bn2.running_mean.uniform_(-0.2, 0.2)
        x.bn2.running_var.uniform_(0.5, 1.5)
        x.bn2.weight.uniform_(0.7, 1.3)
        x.bn2.bias.uniform_(-0.1, 0.1)
        ep = eng.pack_entry(x, torch.float32)
        inp = torch.randn(1, 32, 12, 12)
        want = x.bn2(x.conv2(inp))
        got = F.conv2d(inp, ep.conv2_w.permute(0, 3, 1, 2), ep.conv2_b)
        assert torch.allclose(got, want, atol=1e-5)
        b3 = ep.blocks[2]
        assert b3.skip_w.shape == (728, 256) and b3.seps[1].pw.shape == (728, 728) and b3.seps[0].dw.shape == (3, 3, 256)
        assert b3.start_with_relu and not ep.blocks[0].start_with_relu


def test_input_norm_folding_matches_torch(model):
    """fold_input_norm: conv(w', u8) + b' == conv(w, (u8 / 255 - mean) / std) + b on CPU (the uint8 input stage's
    host arithmetic; the kernel only converts bytes), checked through the oracle's normalise_u8."""
    import importlib
    from helpers import oracle
    eng = importlib.import_module("2023-tifs-istvt_b200.engine")
    O = oracle()
    g = torch.Generator().manual_seed(5)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(32, generator=g) * 0.1
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    u8 = torch.randint(0, 256, (1, 2, 21, 23, 3), generator=g, dtype=torch.uint8)
    want = F.conv2d(O.normalise_u8(u8, mean, std).flatten(0, 1), w, b, stride=2)
    w2, b2 = eng.fold_input_norm(w, b, mean, std)
    got = F.conv2d(u8.flatten(0, 1).permute(0, 3, 1, 2).float(), w2, b2, stride=2)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)


def test_uint8_clip_validation(model):
    with pytest.raises(ValueError, match="uint8 clips must be"):
        model(torch.zeros(1, 6, 3, 300, 300, dtype=torch.uint8))
    with pytest.raises(ValueError, match="CUDA"):
        model(torch.zeros(1, 6, 300, 300, 3, dtype=torch.uint8))


# ------------------------------------------------------------------------------------------------
# ablation transformers: module tree, constructor contract, error behaviour
# ------------------------------------------------------------------------------------------------
def test_ablation_module_trees():
    m = pkg()
    v = m.ViViT(19, 1, 1, 6, depth=1)
    sd = v.state_dict()
    assert sd["pos_embedding"].shape == (1, 6, 362, 728)
    assert sd["space_transformer.layers.0.0.fn.to_qkv.weight"].shape == (1536, 728)
    assert sd["temporal_transformer.layers.0.1.fn.net.0.weight"].shape == (2912, 728)
    assert "temporal_token" in sd and "space_token" in sd and sd["mlp_head.1.weight"].shape == (1, 728)
    w = m.VanillaTr(19, 1, 1, 6, depth=1)
    sd = w.state_dict()
    assert sd["pos_embedding"].shape == (1, 6 * 361 + 1, 728) and sd["cls_token"].shape == (1, 1, 728)
    assert sd["to_patch_embedding.1.weight"].shape == (728, 728)
    assert sd["transformer.layers.0.0.fn.to_out.0.bias"].shape == (728,)
    a = m.TemporalOnlyAttention(728)
    assert a.to_qkv.weight.shape == (1536, 728) and a.scale == 0.125


def test_ablation_constructor_contract():
    m = pkg()
    with pytest.raises(ValueError):
        m.ViViT(19, 1, 1, 6, pool="max")
    assert m.ViViT(19, 1, 1, 6, depth=1, pool="mean").pool == "mean"
    with pytest.raises(ValueError):
        m.VanillaTr(19, 2, 1, 6)                  # 19 % 2
    with pytest.raises(ValueError):
        m.Attention(728, dim_head=32)
    with pytest.raises(ValueError):
        m.XceptionVidTr(variant="joint")


def test_ablation_no_cpu_fallback():
    m = pkg()
    v = m.ViViT(19, 1, 1, 6, depth=1).eval()
    with pytest.raises(ValueError, match="CUDA"):
        v(torch.zeros(1, 6, 728, 19, 19))
    with pytest.raises(ValueError):
        v(torch.zeros(1, 6, 19, 19))
    for blk in (m.Attention(728).eval(), m.TemporalOnlyAttention(728).eval(), m.Transformer(728, 1, 8, 64, 2912).eval()):
        with pytest.raises(ValueError, match="CUDA"):
            blk(torch.zeros(1, 362, 728))
