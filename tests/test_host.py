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


def test_shadow_mode_import_runs_forward():
    """with the package directory itself on sys.path, `from network.models import model_selection`
    must give a model whose forward reaches the device check (ValueError on a CPU tensor), not an ImportError from
    relative imports that climb above a re-rooted `network` package; eval, train and ablation entries all resolve."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {os.path.join(root, '2023-tifs-istvt_b200')!r})\n"
        "from network.models import model_selection\n"
        "import network.vivit.vivit as vv, network.vivit.module as mod, network.xception as xc\n"
        "m = model_selection('resnet_3d', num_out_classes=1)\n"
        "assert type(m) is vv.XceptionVidTr\n"
        "x = torch.zeros(1, 6, 3, 300, 300)\n"
        "for mode in (m.eval(), m.train()):\n"
        "    try:\n"
        "        mode(x)\n"
        "    except ValueError as e:\n"
        "        assert 'CUDA' in str(e), e\n"
        "    else:\n"
        "        raise SystemExit('no ValueError')\n"
        "for ctor in (lambda: vv.XceptionVidTr(variant='vivit'), lambda: mod.Attention(728), lambda: model_selection('xception', 2)):\n"
        "    a = ctor().eval()\n"
        "    inp = x if not isinstance(a, mod.Attention) else torch.zeros(1, 10, 728)\n"
        "    if hasattr(a, 'model') and not hasattr(a, 'vit'): inp = torch.zeros(1, 3, 299, 299)\n"
        "    try:\n"
        "        a(inp)\n"
        "    except ValueError as e:\n"
        "        assert 'CUDA' in str(e), e\n"
        "    else:\n"
        "        raise SystemExit('no ValueError')\n"
        "print('shadow ok')\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and "shadow ok" in r.stdout, r.stdout + r.stderr


def _fake_replicate(model):
    """What torch.nn.parallel.replicate() does to a module tree (torch/nn/parallel/replicate.py), on CPU: shallow
    __dict__ copies, `_parameters` emptied, the parameter copies (non-leaf tensors) set as plain attributes and listed
    in `_former_parameters`."""
    from collections import OrderedDict
    modules = list(model.modules())
    index = {m: i for i, m in enumerate(modules)}
    copies = [m._replicate_for_data_parallel() for m in modules]
    for r in copies:
        r._former_parameters = OrderedDict()
    for i, m in enumerate(modules):
        for k, child in m._modules.items():
            if child is not None:
                setattr(copies[i], k, copies[index[child]])
        for k, p in m._parameters.items():
            if p is not None:
                c = p * 1.0          # non-leaf, like the output of Broadcast.apply
                setattr(copies[i], k, c)
                copies[i]._former_parameters[k] = c
        for k, b in m._buffers.items():
            if b is not None:
                setattr(copies[i], k, b.clone())
    return copies[0]


def test_data_parallel_replica_shares_engine_and_lists_parameters(model):
    """nn.DataParallel (train_CNN.py:185-186): a replica must (a) reach the owner's engine, whose pack cache is
    validated against the OWNER's parameters — the replica's are fresh copies every step — and (b) expose the on-path
    parameter copies, in named_parameters() order, to the training bridge."""
    train = __import__("importlib").import_module("2023-tifs-istvt_b200.train")
    eng_owner = model.engine()
    rep = _fake_replicate(model)
    assert getattr(rep, "_is_replica", False) and len(list(rep.parameters())) == 0
    assert rep._shared is model._shared and rep.engine() is eng_owner
    assert eng_owner._fingerprint(rep) == eng_owner._fingerprint(model)
    rep2 = _fake_replicate(model)                       # next step's replica: new tensors, same fingerprint
    assert eng_owner._fingerprint(rep2) == eng_owner._fingerprint(model)
    before = eng_owner._fingerprint(rep)
    with torch.no_grad():
        model.vit.mlp_head[1].bias.add_(1.0)            # an optimizer step on the owner invalidates it
    assert eng_owner._fingerprint(_fake_replicate(model)) != before
    want = train.on_path_named_parameters(model, True)
    got = train.on_path_named_parameters(rep, True)
    assert [n for n, _ in got] == [n for n, _ in want] and len(got) == 252
    assert all(g.shape == w.shape and not g.is_leaf for (_, g), (_, w) in zip(got, want))
    st = train.FlatState(rep, True, grads_only=True)
    assert st.params is None and st.grads.numel() >= 89467761 and set(st.grad) == {n for n, _ in want}


def test_model_pickles_without_shared_state(model):
    import copy
    m2 = copy.deepcopy(model)
    assert m2._shared is not model._shared and m2._shared.owner() is m2 and m2._engine is None


def test_ablation_paths_refuse_train_mode():
    """train mode means BatchNorm batch statistics + dropout in the reference; the ablation paths fold the
    running statistics, so they must refuse `module.train()` even under torch.no_grad()."""
    m = pkg()
    vivit = m.XceptionVidTr(variant="vivit").train()
    with torch.no_grad(), pytest.raises(NotImplementedError, match="eval"):
        vivit(torch.zeros(1, 6, 3, 300, 300))
    attn = m.Attention(728).train()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        attn(torch.zeros(1, 10, 728))


def test_vanilla_tr_checks_patch_linear_channels():
    """VanillaTr(in_channels != dim) is a valid model (patch Linear, vivit.py:162-167): the channel check must follow
    the Linear's in_features — the error for a wrong channel count names it, the right count passes to the device check."""
    m = pkg()
    v = m.VanillaTr(19, 1, 1, 6, dim=128, depth=1, in_channels=64).eval()
    with pytest.raises(ValueError, match="64"):
        v(torch.zeros(1, 6, 128, 19, 19))
    with pytest.raises(ValueError, match="CUDA"):
        v(torch.zeros(1, 6, 64, 19, 19))


def test_layernorm_fold_algebra():
    """engine.fold_layernorm + the definitions of the two fold GEMMs reproduce Linear(LayerNorm(y)) (module.py:21, :83)."""
    import torch_ops as tops
    eng = __import__("importlib").import_module("2023-tifs-istvt_b200.engine")
    g = torch.Generator().manual_seed(5)
    a = torch.randn(37, 512, generator=g)
    w1, b1 = torch.randn(728, 512, generator=g) * 0.05, torch.randn(728, generator=g)
    gamma, beta = torch.rand(728, generator=g) + 0.5, torch.randn(728, generator=g)
    w2 = torch.randn(1536, 728, generator=g) * 0.04
    stats = torch.zeros(37, 12, 2)
    y = tops.gemm_rowstats(a, w1, b1, stats)
    wf, c, d = eng.fold_layernorm(w2, gamma, beta, torch.float32)
    got = tops.gemm_lnfold(y, wf, tops.ln_stats_finalize(stats, 728), c, d)
    want = F.linear(F.layer_norm(F.linear(a, w1, b1), (728,), gamma, beta), w2)
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()


def test_persisted_pack_round_trip_and_staleness(tmp_path):
    """packed weights written beside a checkpoint come back bit-identical (row-pitched
    tensors at their pitch), and a pack written for other weights is refused."""
    eng = __import__("importlib").import_module("2023-tifs-istvt_b200.engine")
    ops = __import__("importlib").import_module("2023-tifs-istvt_b200.ops")
    torch.manual_seed(3)
    m = pkg().XceptionVidTr().eval()
    path = str(tmp_path / "best.pkl.pack")
    digest = m.save_packed_weights(path)
    assert digest == eng.state_digest(m)
    assert m.load_packed_weights(path)
    got = m.engine()._packs[("cpu", "bf16")]
    want = eng.pack_model(m, torch.bfloat16)

    def leaves(o, out):
        from dataclasses import fields, is_dataclass
        if is_dataclass(o):
            for f in fields(o):
                if f.name != "fingerprint":
                    leaves(getattr(o, f.name), out)
        elif isinstance(o, torch.Tensor):
            out.append(o)
        elif isinstance(o, (list, tuple)):
            for v in o:
                leaves(v, out)
        return out
    a, b = leaves(got, []), leaves(want, [])
    assert len(a) == len(b) > 200
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and x.shape == y.shape and x.stride() == y.stride() and torch.equal(x, y)
    lp = got.layers[0]
    assert lp.w_qkv_f.stride(0) == ops.row_pitch(728) == 768 and lp.w_2.stride(0) == 2912
    assert got.fingerprint == eng._fingerprint(m)
    with torch.no_grad():
        m.vit.transformer.layers[3][2].fn.net[0].bias.add_(1e-3)      # any on-path change invalidates the file
    assert not m.load_packed_weights(path)
