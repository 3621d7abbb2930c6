"""CPU suite (-m "not gpu"): the oracle against the committed golden vectors and against its independent
numpy restatement; the reference's module structure (state_dict naming) and flag surface."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "oracle_small.npz"))


def _load(net, gold, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(prefix)}
    net.load_state_dict(sd)
    return net.eval()


def test_oracle_generator_matches_golden(gold):
    from oracle.networks import define_G
    g = _load(define_G(5, 4, 8, "temporal", 1, 1), gold, "g.")
    with torch.no_grad():
        y = g(torch.from_numpy(gold["x"]))
    assert np.abs(y.numpy() - gold["y"]).max() <= 1e-5
    assert y[:, 3].min() >= 0 and y[:, 3].max() <= 1 and y[:, :3].abs().max() <= 1      # mask sigmoid / rgb tanh


def test_oracle_discriminator_matches_golden(gold):
    from oracle.networks import define_D
    d = _load(define_D(7, 8, 2, "instance", False, 2, True), gold, "d.")
    with torch.no_grad():
        yd = d(torch.from_numpy(gold["xd"]))
    assert len(yd) == 2 and len(yd[0]) == 4                       # num_D scales x (n_layers + 2) feature maps
    assert np.abs(yd[0][-1].numpy() - gold["yd_last0"]).max() <= 1e-5
    assert np.abs(yd[1][-1].numpy() - gold["yd_last1"]).max() <= 1e-5
    assert np.abs(yd[0][0].numpy() - gold["yd_feat00"]).max() <= 1e-5
    # pix2pixHD geometry: 4x4 s2 p2 maps 24 -> 13 -> 7(s2) -> 8(s1) -> 9(s1); second scale starts at 12
    assert yd[0][0].shape[-1] == 13 and yd[0][-1].shape[-1] == 9 and yd[1][0].shape[-1] == 7


def test_oracle_texture_matches_golden_and_numpy(gold):
    from oracle.texture import texture_sample, texture_sample_numpy, composite
    uvp, atlas = torch.from_numpy(gold["uvp"]), torch.from_numpy(gold["atlas"])
    tex, part, texel = texture_sample(uvp, atlas, True)
    assert np.array_equal(part.numpy(), gold["part"]) and np.array_equal(texel.numpy(), gold["texel"])
    assert np.abs(tex.numpy() - gold["tex"]).max() <= 1e-6
    t2, p2, x2 = texture_sample_numpy(gold["uvp"], gold["atlas"], True)      # independent scalar-loop restatement
    assert np.array_equal(p2, gold["part"]) and np.array_equal(x2, gold["texel"])
    assert np.abs(t2 - gold["tex"]).max() <= 1e-5
    comp = composite(torch.from_numpy(gold["fgm"]), torch.from_numpy(gold["bg"]))
    assert np.abs(comp.numpy() - gold["comp"]).max() <= 1e-7


def test_oracle_texture_edge_cases():
    """u,v exactly 0 / 1 / out of range; logit ties -> lowest index; background pixel -> texel (0,0)."""
    from oracle.texture import texture_sample, texture_sample_numpy
    S = 8
    uvp = torch.zeros(1, 73, 2, 3)
    uvp[0, 25:, 0, 0] = -1.0       # u = v = 0
    uvp[0, 25:, 0, 1] = 1.0        # u = v = 1 -> x0 = S-1, x1 clamped
    uvp[0, 25:, 0, 2] = 7.0        # clamps to 1
    uvp[0, 5, 0, 0] = 3.0
    uvp[0, 5, 0, 1] = 3.0
    uvp[0, 9, 0, 2] = 2.0
    uvp[0, 9, 1, 0] = uvp[0, 4, 1, 0] = 2.0      # tie between part 4 and 9 -> 4
    # (1,1): all logits equal -> part 0 (background)
    atlas = torch.rand(24, 3, S, S)
    tex, part, texel = texture_sample(uvp, atlas, True)
    assert part[0, 0, 0] == 5 and tuple(texel[0, 0, 0].tolist()) == (0, 0)
    assert part[0, 0, 1] == 5 and tuple(texel[0, 0, 1].tolist()) == (S - 1, S - 1)
    assert part[0, 0, 2] == 9 and tuple(texel[0, 0, 2].tolist()) == (S - 1, S - 1)
    assert part[0, 1, 0] == 4
    assert part[0, 1, 1] == 0 and tuple(texel[0, 1, 1].tolist()) == (0, 0)
    t2, p2, x2 = texture_sample_numpy(uvp.numpy(), atlas.numpy(), True)
    assert np.array_equal(p2, part.numpy()) and np.array_equal(x2, texel.numpy())
    assert np.abs(t2 - tex.numpy()).max() <= 1e-5
    # without use_mask_texture the blend is renormalised by the foreground probability
    tex_n, _, _ = texture_sample(uvp, atlas, False)
    p0 = torch.softmax(uvp[:, :25], 1)[:, :1]
    assert torch.allclose(tex_n, tex / (1 - p0 + 1e-6), atol=1e-6)


def test_oracle_losses_match_golden(gold):
    from oracle import losses
    uvp = torch.from_numpy(gold["uvp"])
    dp_i, dp_uv = torch.from_numpy(gold["dp_i"]), torch.from_numpy(gold["dp_uv"])
    assert abs(losses.uv_loss(uvp, dp_i, dp_uv).item() - float(gold["l_uv"])) <= 1e-6
    assert abs(losses.prob_loss(uvp, dp_i).item() - float(gold["l_prob"])) <= 1e-6
    # flow warp with zero flow is the identity; constant flow shifts
    img = torch.rand(1, 3, 6, 7)
    assert torch.allclose(losses.flow_warp(img, torch.zeros(1, 2, 6, 7)), img, atol=1e-6)
    fl = torch.zeros(1, 2, 6, 7); fl[:, 0] = 1.0
    assert torch.allclose(losses.flow_warp(img, fl)[..., :-1], img[..., 1:], atol=1e-5)
    assert losses.temporal_loss(img, img, torch.zeros(1, 2, 6, 7)).item() <= 1e-7


def test_oracle_generator_structure_is_pix2pixhd():
    """model.<idx> indices of GlobalGenerator(n_down=2, n_blocks=10): conv at 1, 4, 7; blocks 10..19;
    convT at 20, 23; head conv at 27."""
    from oracle.networks import define_G
    g = define_G(6, 3, 48, "global", 2, 10)
    keys = set(g.state_dict().keys())
    for k in ["model.1.weight", "model.4.weight", "model.7.weight", "model.10.conv_block.1.weight",
              "model.19.conv_block.5.bias", "model.20.weight", "model.23.weight", "model.27.weight"]:
        assert k in keys, k
    assert g.state_dict()["model.20.weight"].shape == (192, 96, 3, 3)      # ConvTranspose2d [Cin, Cout, k, k]
    assert sum(p.numel() for p in g.parameters()) == pytest.approx(7.07e6, rel=0.01)
    x = torch.randn(1, 6, 32, 32)
    with torch.no_grad():
        assert g(x).shape == (1, 3, 32, 32)


def test_product_modules_share_state_dict_with_oracle():
    """Drop-in boundary: same define_G signature, same parameter names and shapes (checkpoint compatible)."""
    from oracle.networks import define_G as oG
    from oracle.pipeline import RenderModel
    from nhvr_b200.networks import define_G
    from nhvr_b200.pipeline import RenderPipeline
    for args in [(9, 4, 48, "temporal", 2, 10), (3, 73, 64, "translate", 2, 5), (3, 3, 48, "bg", 2, 2)]:
        a, b = oG(*args).state_dict(), define_G(*args).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
    kw = dict(size=32, atlas_size=8, ngf_global=8, n_blocks_global=1, ngf_translate=8, n_blocks_translate=1, ngf_bg=8, n_blocks_bg=1)
    a, b = RenderModel(**kw).state_dict(), RenderPipeline(**kw).state_dict()
    assert set(a.keys()) == set(b.keys()) and all(a[k].shape == b[k].shape for k in a)


def test_product_has_no_cpu_fallback():
    from nhvr_b200.networks import define_G
    from nhvr_b200.capi import NhvrError
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    net = define_G(3, 3, 8, "global", 1, 1)
    with pytest.raises(NhvrError):
        with torch.no_grad():
            net(torch.randn(1, 3, 16, 16))


def test_reference_scripts_parse_verbatim():
    """Every flag list of the reference's launch scripts parses with the matching option class."""
    from nhvr_b200.options import TestOptions, TrainOptions, pipeline_kwargs
    flags = json.load(open(os.path.join(GOLD, "ref_flags.json")))
    o = TestOptions().parse(flags["test_start/start.sh"][0]["argv"])
    assert flags["test_start/start.sh"][0]["entry"] == "test.py"
    assert (o.n_downsample_global, o.n_blocks_global, o.ngf_global, o.n_downsample_bg, o.n_blocks_bg) == (2, 10, 48, 2, 2)
    assert o.use_laplace and o.pose_plus_laplace and o.use_mask_texture and o.TexG == "part" and o.which_epoch == "30"
    assert o.pose_nc == 6 and o.loadSize == 512 and o.pose_path == "./keypoints"
    kw = pipeline_kwargs(o)
    assert kw["pose_nc"] == 6 and kw["n_blocks_global"] == 10
    t = TrainOptions().parse(flags["train_start/pretrain_start.sh"][0]["argv"])
    assert (t.lambda_L2, t.lambda_UV, t.lambda_Prob, t.lambda_Temp) == (500, 1000, 10, 500)
    assert t.use_densepose_loss and t.batchSize == 2 and t.which_epoch_TransG == "2" and t.data_ratio == 0.9
    p = TrainOptions().parse(flags["pretrainTrans.sh"][0]["argv"])
    assert p.n_blocks_translate == 5 and p.batchSize == 6 and p.save_epoch_freq == 2 and p.tf_log
    x = TrainOptions().parse(flags["pre_train_tex.sh"][0]["argv"])
    assert x.input_nc == 81 and x.loadSize == 200 and x.lapalce_path and x.gpu_ids == [1]
    with pytest.raises(SystemExit):
        TestOptions().parse(["--gpu_ids", "-1"])
    with pytest.raises(SystemExit):
        TestOptions().parse(["--no_such_flag"])


def test_independent_numpy_restatement_agrees_with_the_torch_oracle(gold):
    """oracle/numpy_ref.py restates ReflectionPad2d / Conv2d / ConvTranspose2d / InstanceNorm2d / the ResnetBlock / the
    odd-size PatchGAN arithmetic (4x4 stride 2 padding 2: 24 -> 13 -> 7 -> 8 -> 9; AvgPool(3,2,1) without padded counts)
    from the operator definitions in plain numpy.  It must reproduce the golden vectors the torch.nn oracle produced
    (tests/golden/oracle_small.npz): the two restatements pin each other."""
    from oracle import numpy_ref as R
    sd = {k: gold[k] for k in gold.files}
    y = R.global_generator(gold["x"], sd, n_down=1, n_blocks=1, final="tanh_sigmoid_last", prefix="g.")
    assert y.shape == gold["y"].shape
    assert np.abs(y - gold["y"]).max() <= 2e-5
    feats = R.multiscale_discriminator(gold["xd"], sd, num_D=2, n_layers=2, prefix="d.")
    assert feats[0][0].shape == gold["yd_feat00"].shape == (1, 8, 13, 13)
    assert feats[0][-1].shape == (1, 1, 9, 9) and feats[1][-1].shape == (1, 1, 6, 6)
    assert np.abs(feats[0][0] - gold["yd_feat00"]).max() <= 2e-5
    assert np.abs(feats[0][-1] - gold["yd_last0"]).max() <= 5e-5
    assert np.abs(feats[1][-1] - gold["yd_last1"]).max() <= 5e-5
    # single operators on odd sizes against their definitions' corner cases
    x = np.arange(2 * 1 * 5 * 7, dtype=np.float64).reshape(2, 1, 5, 7)
    p = R.reflect_pad(x, 2)
    assert p.shape == (2, 1, 9, 11) and p[0, 0, 0, 0] == x[0, 0, 2, 2] and p[0, 0, 8, 10] == x[0, 0, 2, 4]
    a = R.avgpool3s2(x)
    assert a.shape == (2, 1, 3, 4) and a[0, 0, 0, 0] == x[0, 0, :2, :2].mean() and a[0, 0, 2, 3] == x[0, 0, 3:5, 5:7].mean()


def test_vgg19_oracle_and_module_share_torchvision_parameter_names():
    """The perceptual-loss stack: five taps at /1, /2, /4, /8, /16 resolution; the B200 module holds the same ``features.<idx>`` keys
    (a torchvision vgg19 checkpoint loads into both)."""
    import torch
    from oracle.networks import Vgg19
    from oracle.losses import vgg_loss
    from nhvr_b200.networks import Vgg19B200
    torch.manual_seed(0)
    ref = Vgg19().eval()
    x = torch.rand(1, 3, 32, 48)
    feats = ref(x)
    assert [tuple(f.shape[1:]) for f in feats] == [(64, 32, 48), (128, 16, 24), (256, 8, 12), (512, 4, 6), (512, 2, 3)]
    net = Vgg19B200()
    assert sorted(net.state_dict().keys()) == sorted(ref.state_dict().keys())
    net.load_state_dict(ref.state_dict())
    assert not any(p.requires_grad for p in net.parameters())
    assert float(vgg_loss(ref, x, x)) == 0.0 and float(vgg_loss(ref, x, torch.rand(1, 3, 32, 48))) > 0.0
