"""GPU parity: the sm_100a path (through the C-ABI) against the fp32 oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): integer part / texel indices bit-exact; rendered frames max-abs
<= 2e-2 on [-1,1] and PSNR >= 45 dB.  Intermediate bf16 conv chains are compared at the same bound.
"""
import math

import os

import pytest
import torch

pytestmark = pytest.mark.gpu

# parity bars: fp16 operands (the default and the only operand type a parity claim is made for) meet BASELINE's 2e-2 / 45 dB.
# bf16 operands (NHVR_OPERAND=bf16, kept because north_star names bf16) are measured ~4x outside the bar at full depth
# (profiles/r01_precision.log); their entry is a regression guard for that alternative path, NOT a north_star bar
BARS = {"f16": (2e-2, 45.0), "bf16": (1.5e-1, 38.0)}


@pytest.fixture(params=["f16", "bf16"])
def operand(request, cuda_dev):
    from nhvr_b200 import capi
    prev = capi.operand_dtype()
    capi.set_operand_dtype(request.param)
    yield request.param
    capi.set_operand_dtype(prev)


def smooth_atlas(C, S, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(24, C, 6, 6, generator=g)
    return torch.tanh(torch.nn.functional.interpolate(low, size=(S, S), mode="bicubic", align_corners=False)).to(dev)


def psnr(a, b, peak=2.0):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * math.log10(peak * peak / mse)


def _pair_G(dev, *args, seed=0):
    from nhvr_b200.networks import define_G
    from oracle.networks import define_G as oracle_define_G
    torch.manual_seed(seed)
    ref = oracle_define_G(*args).to(dev).eval()
    net = define_G(*args)
    net.load_state_dict(ref.state_dict())
    return net, ref


@pytest.mark.parametrize("netG,cin,cout,ngf,nd,nb,size,batch", [
    ("temporal", 9, 4, 16, 2, 2, 64, 2),
    ("temporal", 9, 4, 48, 2, 10, 128, 1),
    ("translate", 3, 73, 32, 2, 3, 96, 1),
    ("bg", 3, 3, 48, 2, 2, 72, 1),
    ("global", 6, 3, 24, 1, 1, 40, 3),
])
def test_generator_parity(cuda_dev, operand, netG, cin, cout, ngf, nd, nb, size, batch):
    net, ref = _pair_G(cuda_dev, cin, cout, ngf, netG, nd, nb)
    torch.manual_seed(1)
    x = torch.rand(batch, cin, size, size, device=cuda_dev) * 2 - 1          # image-range inputs
    with torch.no_grad():
        y = net(x)
        y_ref = ref(x)
    assert y.shape == y_ref.shape
    tol, db = BARS[operand]
    err = (y - y_ref).abs().max().item()
    scale = max(1.0, y_ref.abs().max().item())
    assert err <= tol * scale, (operand, err, scale)
    if netG != "translate":
        assert psnr(y, y_ref) >= db


def test_generator_rect_and_repack(cuda_dev):
    """non-square input; weights changed in place must be re-packed (optimizer-step semantics)."""
    net, ref = _pair_G(cuda_dev, 5, 3, 16, "global", 2, 1)
    x = torch.rand(1, 5, 48, 80, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        assert (net(x) - ref(x)).abs().max().item() <= 2e-2
        y0 = net(x).clone()
        for p, q in zip(net.parameters(), ref.parameters()):
            p.mul_(1.5)
            q.mul_(1.5)
        y1 = net(x)
        assert (y1 - ref(x)).abs().max().item() <= 2e-2
        assert (y1 - y0).abs().max().item() > 1e-3                   # the new weights were really used


def test_generator_rejects_cpu(cuda_dev):
    from nhvr_b200.capi import NhvrError
    net, _ = _pair_G(cuda_dev, 3, 3, 16, "global", 1, 1)
    with pytest.raises(NhvrError):
        with torch.no_grad():
            net(torch.randn(1, 3, 32, 32))


@pytest.mark.parametrize("N,H,W,S,Ctex,mask", [(1, 64, 64, 200, 3, True), (2, 33, 47, 50, 3, False), (1, 40, 56, 64, 18, True)])
def test_texture_sample_parity(cuda_dev, N, H, W, S, Ctex, mask):
    from nhvr_b200 import ops
    from oracle.texture import texture_sample
    torch.manual_seed(3)
    uvp = torch.randn(N, 73, H, W, device=cuda_dev) * 2.0
    uvp[:, 25:] *= 0.7
    # edge cases: exact 0 / 1 UV, logit ties
    uvp[0, 25:, 0, 0] = -1.0
    uvp[0, 25:, 0, 1] = 1.0
    uvp[0, 25:, 0, 2] = 5.0
    uvp[0, :25, 1, 0] = 0.25
    uvp[0, 3, 1, 1] = uvp[0, 7, 1, 1] = 9.0
    atlas = torch.rand(24, Ctex, S, S, device=cuda_dev) * 2 - 1
    tex, part, texel = ops.texture_sample(uvp, ops.atlas_to_channels_last(atlas), Ctex, mask)
    tex_r, part_r, texel_r = texture_sample(uvp, atlas, mask)
    assert torch.equal(part, part_r)                      # bit-exact integer contract
    assert torch.equal(texel, texel_r)
    assert (tex - tex_r).abs().max().item() <= 1e-4 * max(1.0, tex_r.abs().max().item())
    assert part[0, 1, 0].item() == 0 and part[0, 1, 1].item() == 3      # ties: lowest index wins



def _conv_case(dev, kind, cin, cout, k, stride, pad, N, H, W, halo, epi, act, env=None, tap_pairing=False):
    """One conv through the C-ABI (pack -> plan -> conv) against torch.nn.functional on the 16-bit-rounded operands.
    Returns (max abs err / max|ref|, statistics rel err, plan info)."""
    import torch.nn.functional as F
    from nhvr_b200 import ops, capi
    old = {}
    for kk, vv in (env or {}).items():
        old[kk] = os.environ.get(kk)
        os.environ[kk] = vv
    try:
        plan = ops.ConvPlan(kind, cin, cout, k, stride, pad, N, H, W, halo, epi, act, allow_tap_pairing=tap_pairing)
    finally:
        for kk, vv in old.items():
            if vv is None:
                os.environ.pop(kk, None)
            else:
                os.environ[kk] = vv
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + k)
    x = (torch.rand(N, cin, H, W, generator=g) * 2 - 1).to(dev)
    transposed = kind == capi.CONV_TRANSPOSE
    w = (torch.randn(*((cin, cout, k, k) if transposed else (cout, cin, k, k)), generator=g) * (1.0 / (cin * k * k) ** 0.5)).to(dev)
    b = (torch.rand(cout, generator=g) - 0.5).to(dev)
    rnd = (lambda t: t.half().float()) if capi.operand_dtype() == "f16" else (lambda t: t.bfloat16().float())
    xin = ops.P8Buffer(plan.in_desc.copy(), dev)
    ops.pack_nchw([x], xin)
    plan.pack_weights(w)
    xr, wr = rnd(x), rnd(w)
    if transposed:
        ref = F.conv_transpose2d(xr, wr, stride=2, padding=pad, output_padding=1 if k == 3 else 0)
    else:
        xp = F.pad(xr, (pad,) * 4, mode="reflect" if halo == capi.HALO_REFLECT else "constant")
        ref = F.conv2d(xp, wr, stride=stride)
    assert (plan.Ho, plan.Wo) == tuple(ref.shape[-2:])
    stat_err = 0.0
    if epi == capi.EPI_RAW_STATS:
        raw = ops.P8Buffer(plan.raw_desc(), dev)
        stats = torch.zeros(N * plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
        plan.forward(xin, raw.ptr, stats=stats)
        out = ops.unpack_nchw(raw, cout)
        st = stats.view(N, plan.Cout8 * 8, 4)[:, :cout, :2].double()
        s_ref = torch.stack([ref.double().sum((2, 3)), (ref.double() ** 2).sum((2, 3))], -1)
        stat_err = ((st - s_ref).abs() / (1.0 + s_ref.abs())).max().item()
    else:
        out = torch.empty(N, cout, plan.Ho, plan.Wo, dtype=torch.float32, device=dev)
        plan.forward(xin, out.data_ptr(), bias=b)
        ref = ref + b.view(1, -1, 1, 1)
        if act == capi.ACT_TANH:
            ref = torch.tanh(ref)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    return err, stat_err, plan.info()


@pytest.mark.parametrize("name,env,kind,cin,cout,k,stride,pad,N,H,W,halo,epi,act,expect", [
    # CTA pairs (cta_group::2): odd tile count per image, two images, statistics epilogue
    ("pair_192", None, "CONV", 192, 192, 3, 1, 1, 2, 35, 37, "R", "RAW_STATS", "NONE", dict(Npad=192)),
    ("pair_forced_256", {"NHVR_CONV_PAIR": "1"}, "CONV", 64, 256, 3, 1, 1, 1, 20, 50, "R", "RAW_STATS", "NONE", dict(Npad=256)),
    ("pair_s2_192", None, "CONV", 96, 192, 3, 2, 1, 1, 66, 70, "Z", "RAW_STATS", "NONE", dict(Npad=192)),
    ("pair_off_192", {"NHVR_CONV_PAIR": "0"}, "CONV", 192, 192, 3, 1, 1, 1, 35, 37, "R", "RAW_STATS", "NONE", dict(Npad=192)),
    # M replication, stacked rows (W multiple of 128) and consecutive positions (W = 100), fp32 head epilogue
    ("mrep_stacked_7x7", None, "CONV", 16, 73, 7, 1, 3, 2, 224, 128, "R", "BIAS_ACT_F32", "NONE", dict(slab_min=1000)),
    ("mrep_linear_3x3", None, "CONV", 32, 48, 3, 1, 1, 2, 200, 200, "R", "RAW_STATS", "NONE", dict(slab_min=900)),
    ("mrep_s2", None, "CONV", 48, 96, 3, 2, 1, 3, 256, 256, "Z", "RAW_STATS", "NONE", dict(slab_min=1200)),
    ("mrep_forced_3", {"NHVR_CONV_MREP": "3"}, "CONV", 16, 64, 3, 1, 1, 1, 31, 128, "R", "BIAS_ACT_F32", "TANH", None),
    # transposed conv with the N split (4 accumulators x 64 columns), row mode head
    ("convT_split", None, "CONV_TRANSPOSE", 64, 128, 3, 2, 1, 1, 24, 40, "Z", "RAW_STATS", "NONE", dict(nsplit=2)),
    ("rowmode_head", None, "CONV", 48, 4, 7, 1, 3, 1, 40, 150, "R", "BIAS_ACT_F32", "TANH", dict(njobs=7)),
    # row mode with 4 stacked output rows per tile (shuffle epilogue): ragged right edge (300 = 2 x 122 + 56), 130 rows = 32 x 4 + 2
    ("rowmode_head_mrep", None, "CONV", 48, 4, 7, 1, 3, 2, 130, 300, "R", "BIAS_ACT_F32", "TANH", dict(njobs=7, slab_min=1200)),
    ("rowmode_5x5_3out", None, "CONV", 24, 3, 5, 1, 2, 3, 67, 140, "Z", "BIAS_ACT_F32", "NONE", dict(njobs=5)),
])
def test_conv_lowering_variants(cuda_dev, name, env, kind, cin, cout, k, stride, pad, N, H, W, halo, epi, act, expect):
    """Every lowering of the shift-GEMM kernel (CTA pair / M replication / N split / row mode) at a size where the
    plan builder really picks it, against torch on identically rounded operands: the only differences left are the
    fp32 accumulation order and the 16-bit rounding of the RAW output (<= 2^-11 relative)."""
    from nhvr_b200 import capi
    err, stat_err, info = _conv_case(cuda_dev, getattr(capi, kind), cin, cout, k, stride, pad, N, H, W,
                                     capi.HALO_REFLECT if halo == "R" else capi.HALO_ZERO, getattr(capi, "EPI_" + epi),
                                     getattr(capi, "ACT_" + act), env)
    for key, val in (expect or {}).items():
        if key == "slab_min":
            assert info["slab_units"] >= val, (name, info)
        else:
            assert info[key] == val, (name, info)
    tol = 2e-3 if epi == "RAW_STATS" else 2e-4
    assert err <= tol, (name, err, info)
    assert stat_err <= 2e-3, (name, stat_err)



@pytest.mark.parametrize("cin,cout,k,pad,N,H,W,halo,epi", [
    (3, 64, 7, 3, 2, 224, 128, "R", "RAW_STATS"),        # the pose stem: odd kw (the 8th column is a zero weight), M-replicated
    (3, 64, 7, 3, 1, 40, 72, "R", "RAW_STATS"),          # same, small: one M block
    (6, 24, 4, 2, 1, 33, 41, "Z", "BIAS_ACT_F32"),       # even kw, zero padding
    (8, 16, 3, 1, 2, 20, 150, "R", "RAW_STATS"),         # all eight channel slots used
])
def test_conv_tap_pairing(cuda_dev, cin, cout, k, pad, N, H, W, halo, epi):
    """Single-plane tap-paired input format of the narrow-input stems (K group 1 = the next pixel): same result as
    the two-plane lowering, against torch on identically rounded operands."""
    from nhvr_b200 import capi
    args = (cuda_dev, capi.CONV, cin, cout, k, 1, pad, N, H, W, capi.HALO_REFLECT if halo == "R" else capi.HALO_ZERO,
            getattr(capi, "EPI_" + epi), capi.ACT_NONE)
    err, stat_err, info = _conv_case(*args, tap_pairing=True)
    assert info["kcp"] == 1 and info["njobs"] == k * ((k + 1) // 2), info
    assert err <= (2e-3 if epi == "RAW_STATS" else 2e-4) and stat_err <= 2e-3, (err, stat_err, info)
    err2, _, info2 = _conv_case(*args, tap_pairing=False)
    assert info2["kcp"] == 2 and err2 <= (2e-3 if epi == "RAW_STATS" else 2e-4)


def test_texture_sample_matches_numpy_loops(cuda_dev):
    from nhvr_b200 import ops
    from oracle.texture import texture_sample_numpy
    torch.manual_seed(5)
    uvp = torch.randn(1, 73, 6, 7, device=cuda_dev)
    atlas = torch.rand(24, 3, 16, 16, device=cuda_dev)
    tex, part, texel = ops.texture_sample(uvp, ops.atlas_to_channels_last(atlas), 3, True)
    t2, p2, x2 = texture_sample_numpy(uvp.cpu().numpy(), atlas.cpu().numpy(), True)
    assert (part.cpu().numpy() == p2).all() and (texel.cpu().numpy() == x2).all()
    assert abs(tex.cpu().numpy() - t2).max() <= 1e-5


@pytest.mark.parametrize("N,H,W,batched", [(1, 64, 64, False), (3, 30, 50, False), (2, 31, 33, True), (4, 512, 512, False)])
def test_composite_parity(cuda_dev, N, H, W, batched):
    from nhvr_b200 import ops
    from oracle.texture import composite
    torch.manual_seed(7)
    fgm = torch.rand(N, 4, H, W, device=cuda_dev)
    bg = torch.rand(*((N, 3, H, W) if batched else (3, H, W)), device=cuda_dev)
    out = ops.composite(fgm, bg)
    ref = composite(fgm, bg)
    assert (out - ref).abs().max().item() <= 1e-6


def test_pack_apply_unpack_roundtrip(cuda_dev):
    """pack -> unpack is bf16 rounding; IN-apply matches InstanceNorm2d+ReLU+ReflectionPad2d."""
    from nhvr_b200 import ops, capi
    torch.manual_seed(9)
    x = torch.randn(2, 11, 13, 17, device=cuda_dev)
    for split, halo, pad in [(0, capi.HALO_REFLECT, (3, 3, 3, 3)), (1, capi.HALO_ZERO, (1, 1, 1, 1)), (0, capi.HALO_ZERO, (0, 0, 1, 1))]:
        buf = ops.P8Buffer(ops.make_desc(2, 2, 13, 17, pad, split, halo))
        ops.pack_nchw([x[:, :4], x[:, 4:]], buf)
        y = ops.unpack_nchw(buf, 11)
        want = x.half().float() if capi.operand_dtype() == "f16" else x.bfloat16().float()
        assert torch.equal(y, want)


def _small_kw():
    return dict(pose_nc=3, tex_nc=3, size=64, atlas_size=32, ngf_global=16, n_downsample_global=2, n_blocks_global=2,
                ngf_translate=16, n_downsample_translate=2, n_blocks_translate=1, ngf_bg=16, n_downsample_bg=2, n_blocks_bg=1)


def _pair_pipeline(dev, kw, seed=11):
    from nhvr_b200.pipeline import RenderPipeline
    from oracle.pipeline import RenderModel
    torch.manual_seed(seed)
    ref = RenderModel(**kw).to(dev).eval()
    with torch.no_grad():
        ref.atlas.copy_(smooth_atlas(kw["tex_nc"], kw["atlas_size"], dev))     # a texture, not white noise
        ref.bg.copy_(torch.tanh(torch.nn.functional.interpolate(torch.randn(1, 3, 4, 4, device=dev), size=kw["size"], mode="bicubic"))[0])
    pipe = RenderPipeline(**kw).to(dev)
    pipe.load_state_dict(ref.state_dict())
    return pipe, ref


def test_pipeline_stage_parity(cuda_dev):
    """Every stage of the path against the oracle ON THE SAME INPUTS (teacher-forced), 3 temporal steps:
    UV generator within 2e-2 of its output scale, integer part/texel bit-exact, texture 1e-4, generator
    frame and composite within 2e-2 / 45 dB."""
    from nhvr_b200 import ops
    from oracle.texture import texture_sample, composite
    pipe, ref = _pair_pipeline(cuda_dev, _small_kw())
    poses = torch.rand(3, 3, 64, 64, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        bg_r = ref.refine_bg()
        bg = pipe.refine_bg()
        assert (bg - bg_r).abs().max().item() <= 2e-2
        prev = torch.zeros(1, 3, 64, 64, device=cuda_dev)
        for t in range(3):
            o = ref.render_frame(poses[t:t + 1], prev, bg_r)
            uvp = pipe.netTransG(poses[t:t + 1])
            assert (uvp - o["uvp"]).abs().max().item() <= 2e-2 * max(1.0, o["uvp"].abs().max().item())
            tex, part, texel = ops.texture_sample(o["uvp"].contiguous(), pipe.atlas_channels_last(), 3, True)
            assert torch.equal(part, o["part"]) and torch.equal(texel, o["texel"])          # bit-exact
            assert (tex - o["tex"]).abs().max().item() <= 1e-4
            fgm = pipe.netG(o["tex"].contiguous(), poses[t:t + 1], prev)
            assert (fgm - o["fgm"]).abs().max().item() <= 2e-2 and psnr(fgm, o["fgm"]) >= 45.0
            out = ops.composite(o["fgm"].contiguous(), bg_r.contiguous())
            assert (out - o["out"]).abs().max().item() <= 1e-6
            prev = o["out"]



def test_pipeline_1024_laplace_stage_parity(cuda_dev):
    """BASELINE configs[4] shape: 1024 x 1024 with the 6-channel pose input (2D + LaplaceProj) and background refinement;
    one teacher-forced step of every stage against the oracle on the same inputs (width 1030 rows: M replication is
    capped by the buffer slack, 4 x 128-pixel segments per row plus a remainder)."""
    from nhvr_b200 import ops
    kw = dict(pose_nc=6, tex_nc=3, size=1024, atlas_size=64, ngf_global=16, n_downsample_global=2, n_blocks_global=1,
              ngf_translate=16, n_downsample_translate=2, n_blocks_translate=1, ngf_bg=16, n_downsample_bg=2, n_blocks_bg=1)
    pipe, ref = _pair_pipeline(cuda_dev, kw, seed=23)
    torch.manual_seed(5)
    pose = torch.tanh(torch.nn.functional.interpolate(torch.randn(1, 6, 32, 32, device=cuda_dev), size=1024, mode="bilinear") * 2)
    prev = torch.rand(1, 3, 1024, 1024, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        bg_r = ref.refine_bg()
        assert (pipe.refine_bg() - bg_r).abs().max().item() <= 2e-2
        o = ref.render_frame(pose, prev, bg_r)
        uvp = pipe.netTransG(pose)
        assert (uvp - o["uvp"]).abs().max().item() <= 2e-2 * max(1.0, o["uvp"].abs().max().item())
        tex, part, texel = ops.texture_sample(o["uvp"].contiguous(), pipe.atlas_channels_last(), 3, True)
        assert torch.equal(part, o["part"]) and torch.equal(texel, o["texel"])
        assert (tex - o["tex"]).abs().max().item() <= 1e-4
        fgm = pipe.netG(o["tex"].contiguous(), pose, prev)
        assert (fgm - o["fgm"]).abs().max().item() <= 2e-2 and psnr(fgm, o["fgm"]) >= 45.0


def test_pipeline_free_running(cuda_dev):
    """Whole path free-running (its own UV-generator output feeds the lookup, its own frames feed back),
    eager and CUDA-graph, small configuration, default ("strict", split precision) mode: north_star's 2e-2 / 45 dB over
    3 frames (the full-size configuration and the recurrence's own error growth: tests/test_gpu_split3.py)."""
    pipe, ref = _pair_pipeline(cuda_dev, _small_kw())
    poses = torch.rand(3, 3, 64, 64, device=cuda_dev) * 2 - 1
    frames_ref = ref.render_clip(poses)
    outs = []
    for use_graph in (False, True):
        frames = pipe.render_clip(poses, use_graph=use_graph)
        outs.append(frames)
        assert psnr(frames, frames_ref) >= 45.0, (use_graph, psnr(frames, frames_ref))
        assert (frames - frames_ref).abs().max().item() <= 2e-2, (use_graph, (frames - frames_ref).abs().max().item())
    assert (outs[0] - outs[1]).abs().max().item() <= 5e-3          # graph replay == eager up to atomic-order noise


def test_pipeline_lockstep_clips_equal_single(cuda_dev):
    """B clips advanced as a batch give the same frames as each clip rendered alone (sharding is
    idempotent; differences are fp32-atomic statistics ordering only)."""
    kw = _small_kw()
    kw.update(n_downsample_global=1, n_blocks_global=1, n_downsample_translate=1, n_downsample_bg=1)
    pipe, _ = _pair_pipeline(cuda_dev, kw, seed=13)
    poses = torch.rand(2, 3, 3, 64, 64, device=cuda_dev) * 2 - 1
    both = pipe.render_clips(poses)
    for b in range(2):
        single = pipe.render_clip(poses[b])
        assert (both[b] - single).abs().max().item() <= 5e-3


@pytest.mark.parametrize("pinned", [True, False])
def test_pipeline_host_clips_stream_under_compute(cuda_dev, pinned):
    """Host-resident poses in / frames out (copies double-buffered on side streams) equal the
    device-resident call; 5 steps so both staging slots are reused."""
    kw = _small_kw()
    kw.update(n_downsample_global=1, n_blocks_global=1, n_downsample_translate=1, n_downsample_bg=1)
    pipe, _ = _pair_pipeline(cuda_dev, kw, seed=17)
    poses_h = torch.rand(2, 5, 3, 64, 64) * 2 - 1
    out_h = torch.full((2, 5, 3, 64, 64), float("nan"))
    if pinned:
        poses_h, out_h = poses_h.pin_memory(), out_h.pin_memory()
    dev_frames = pipe.render_clips(poses_h.to(cuda_dev))
    for _ in range(2):
        ret = pipe.render_clips(poses_h, out=out_h)
        assert ret is out_h and torch.isfinite(out_h).all()
        assert (out_h - dev_frames.cpu()).abs().max().item() <= 5e-3
        out_h.fill_(float("nan"))


# ------------------------------------------------------------------ training side: discriminator + losses (forward)
@pytest.mark.parametrize("getIntermFeat,num_D,size,batch", [(True, 2, 96, 2), (False, 2, 64, 1), (True, 3, 128, 1)])
def test_discriminator_parity(cuda_dev, getIntermFeat, num_D, size, batch):
    from nhvr_b200.networks import define_D
    from oracle.networks import define_D as oracle_define_D
    torch.manual_seed(21)
    ref = oracle_define_D(6, 32, 3, "instance", False, num_D, getIntermFeat).to(cuda_dev).eval()
    net = define_D(6, 32, 3, "instance", False, num_D, getIntermFeat)
    net.load_state_dict(ref.state_dict())
    x = torch.rand(batch, 6, size, size, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        y, y_ref = net(x), ref(x)
    assert len(y) == len(y_ref) == num_D
    for a_scale, b_scale in zip(y, y_ref):
        assert len(a_scale) == len(b_scale)
        for a, b in zip(a_scale, b_scale):
            assert a.shape == b.shape                      # odd PatchGAN sizes (size/2+1, ...) included
            assert (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item())


def test_losses_parity(cuda_dev):
    """fp32 losses within 1e-3 relative of the oracle (BASELINE.json north_star)."""
    from nhvr_b200 import losses as L
    from oracle import losses as O
    torch.manual_seed(23)
    dev = cuda_dev
    a, b = torch.rand(2, 3, 70, 90, device=dev) * 2 - 1, torch.rand(2, 3, 70, 90, device=dev) * 2 - 1

    def close(x, y):
        return abs(float(x) - float(y)) <= 1e-3 * max(abs(float(y)), 1e-6)
    assert close(L.mse(a, b), O.l2_loss(a, b))
    assert close(L.l1(a, b), torch.nn.functional.l1_loss(a, b))
    preds = [[torch.randn(2, 8, 20, 20, device=dev), torch.randn(2, 1, 11, 11, device=dev)],
             [torch.randn(2, 8, 10, 10, device=dev), torch.randn(2, 1, 6, 6, device=dev)]]
    preds2 = [[torch.randn_like(t) for t in s] for s in preds]
    for real in (True, False):
        assert close(L.gan_loss(preds, real), O.gan_loss(preds, real))
    assert close(L.feature_matching_loss(preds, preds2, 0, 2), O.feature_matching_loss(preds, preds2, 0, 2))
    uvp = torch.randn(2, 73, 33, 47, device=dev) * 2
    dp_i = torch.randint(0, 25, (2, 33, 47), device=dev)
    dp_uv = torch.rand(2, 2, 33, 47, device=dev)
    uv, prob = L.uv_prob_losses(uvp, dp_i, dp_uv)
    assert close(uv, O.uv_loss(uvp, dp_i, dp_uv)) and close(prob, O.prob_loss(uvp, dp_i))
    uv0, _ = L.uv_prob_losses(uvp, torch.zeros_like(dp_i), dp_uv)          # no foreground at all -> 0, not NaN
    assert float(uv0) == 0.0
    flow = torch.randn(2, 2, 70, 90, device=dev) * 3
    assert close(L.temporal_loss(a, b, flow), O.temporal_loss(a, b, flow))
    assert close(L.temporal_loss(a, a, torch.zeros_like(flow)), 0.0) or float(L.temporal_loss(a, a, torch.zeros_like(flow))) < 1e-7


# ------------------------------------------------------------------ backward (dgrad / wgrad / IN backward) vs torch autograd
@pytest.mark.parametrize("netG,cin,cout,ngf,nd,nb,size,batch", [
    ("global", 5, 3, 16, 1, 1, 48, 2),
    ("temporal", 9, 4, 16, 2, 2, 64, 1),
    ("translate", 3, 73, 16, 2, 1, 64, 2),
])
def test_generator_backward_parity(cuda_dev, netG, cin, cout, ngf, nd, nb, size, batch):
    """Parameter and input gradients of the conv chain against torch autograd on the fp32 oracle.  The backward
    is exact for the forward that was executed; against the fp32 oracle the difference is dominated by ReLU
    masks that flip where |z| is below the 16-bit rounding of the stored conv output (a fraction f of the
    pixels gives a relative gradient error of sqrt(f)); measured on the full 26-conv generator: cosine 0.996-1.0,
    max error 6-14 % of each tensor's max (profiles/r01_grad_probe.log).  Bounds: cosine >= 0.995, relative L2 <= 10 %."""
    net, ref = _pair_G(cuda_dev, cin, cout, ngf, netG, nd, nb, seed=31)
    torch.manual_seed(32)
    x = (torch.rand(batch, cin, size, size, device=cuda_dev) * 2 - 1).requires_grad_(True)
    x_ref = x.detach().clone().requires_grad_(True)
    wgt = torch.randn(batch, cout, size, size, device=cuda_dev)
    y = net(x)
    assert y.requires_grad
    (y * wgt).mean().backward()
    (ref(x_ref) * wgt).mean().backward()
    pairs = [("input", x.grad, x_ref.grad)]
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        pairs.append((k, p.grad, q.grad))
    for name, a, b in pairs:
        assert a is not None and a.shape == b.shape, name
        scale = b.abs().max().item()
        if name.endswith(".bias") and scale < 1e-9:        # bias in front of an affine-free IN: exactly zero
            assert a.abs().max().item() == 0.0
            continue
        err = (a - b).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
        if name.endswith(".bias") and name != list(dict(net.named_parameters()))[-1]:
            continue                                         # oracle's IN-cancelled bias grads are round-off noise
        rel_l2 = ((a - b).double().norm() / b.double().norm().clamp(min=1e-30)).item()
        assert rel_l2 <= 1.0e-1, (name, rel_l2)
        assert cos >= 0.995, (name, cos)
        assert err <= 5e-1 * scale + 1e-12, (name, err, scale)     # sparse mask flips: loose on the max norm


def test_generator_two_forwards_one_backward(cuda_dev):
    """Two forward calls of one module before backward (two frames of a step) keep separate saved state."""
    net, ref = _pair_G(cuda_dev, 3, 3, 16, "global", 1, 1, seed=41)
    a = torch.rand(1, 3, 32, 32, device=cuda_dev) * 2 - 1
    b = torch.rand(1, 3, 32, 32, device=cuda_dev) * 2 - 1
    (net(a).mean() + 2 * net(b).mean()).backward()
    (ref(a).mean() + 2 * ref(b).mean()).backward()
    p, q = net.model[1].weight.grad, ref.model[1].weight.grad
    assert (p - q).abs().max().item() <= 2.5e-1 * q.abs().max().item()


def test_uv_pretrain_objective_and_step(cuda_dev):
    """configs[1] in miniature: UV-generator forward, lambda_UV/lambda_Prob objective, backward, Adam step —
    loss and its gradient w.r.t. the network output against torch autograd; the loss goes down."""
    from nhvr_b200 import losses as L
    from oracle import losses as O
    net, ref = _pair_G(cuda_dev, 3, 73, 16, "translate", 2, 1, seed=51)
    torch.manual_seed(52)
    pose = torch.rand(2, 3, 64, 64, device=cuda_dev) * 2 - 1
    dp_i = torch.randint(0, 25, (2, 64, 64), device=cuda_dev)
    dp_uv = torch.rand(2, 2, 64, 64, device=cuda_dev)
    uvp = (torch.randn(2, 73, 64, 64, device=cuda_dev)).requires_grad_(True)
    uvp_r = uvp.detach().clone().requires_grad_(True)
    loss = L.uv_prob_objective(uvp, dp_i, dp_uv, 1000.0, 10.0)
    loss_r = 1000.0 * O.uv_loss(uvp_r, dp_i, dp_uv) + 10.0 * O.prob_loss(uvp_r, dp_i)
    assert abs(loss.item() - loss_r.item()) <= 1e-3 * abs(loss_r.item())
    loss.backward(); loss_r.backward()
    assert (uvp.grad - uvp_r.grad).abs().max().item() <= 1e-4 * uvp_r.grad.abs().max().item() + 1e-9
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    vals = []
    for _ in range(6):
        opt.zero_grad()
        l = L.uv_prob_objective(net(pose), dp_i, dp_uv, 1000.0, 10.0)
        l.backward()
        opt.step()
        vals.append(l.item())
    assert vals[-1] < vals[0], vals


def test_pipeline_backward_parity(cuda_dev):
    """One differentiable frame of the whole path (UV generator -> lookup -> temporal generator -> bg net ->
    composite) with the G-side objective lambda_L2*L2 + lambda_UV*UV + lambda_Prob*Prob: gradients of every
    parameter group (three networks, atlas, background image) against torch autograd on the oracle."""
    from nhvr_b200 import losses as L
    from oracle import losses as O
    from oracle.texture import texture_sample, composite
    pipe, ref = _pair_pipeline(cuda_dev, _small_kw(), seed=61)
    torch.manual_seed(62)
    pose = torch.rand(2, 3, 64, 64, device=cuda_dev) * 2 - 1
    prev = torch.rand(2, 3, 64, 64, device=cuda_dev) * 2 - 1
    real = torch.rand(2, 3, 64, 64, device=cuda_dev) * 2 - 1
    dp_i = torch.randint(0, 25, (2, 64, 64), device=cuda_dev)
    dp_uv = torch.rand(2, 2, 64, 64, device=cuda_dev)
    r = pipe.forward_train(pose, prev)
    loss = L.mse_diff(r["out"], real, 500.0) + L.uv_prob_objective(r["uvp"], dp_i, dp_uv, 1000.0, 10.0)
    loss.backward()
    uvp = ref.netTransG(pose)
    tex, _, _ = texture_sample(uvp, ref.atlas, True)
    fgm = ref.netG(torch.cat([tex, pose, prev], 1))
    out = composite(fgm, ref.netBG(ref.bg.unsqueeze(0))[0])
    loss_r = 500.0 * O.l2_loss(out, real) + 1000.0 * O.uv_loss(uvp, dp_i, dp_uv) + 10.0 * O.prob_loss(uvp, dp_i)
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) <= 5e-3 * abs(loss_r.item())
    checked = 0
    for (k, p), (_, q) in zip(sorted(pipe.named_parameters()), sorted(ref.named_parameters())):
        assert p.grad is not None, k
        if k.endswith(".bias") and p.grad.abs().max().item() == 0.0:
            continue            # bias in front of an affine-free IN: exactly zero here, round-off noise in the oracle
        cos = torch.nn.functional.cosine_similarity(p.grad.flatten().double(), q.grad.flatten().double(), dim=0).item()
        assert cos >= 0.99, (k, cos)
        checked += 1
    assert checked >= 20


def test_sampler_and_composite_backward_exact(cuda_dev):
    """The two memory-bound stages' gradients against torch autograd on identical fp32 inputs (tight bounds)."""
    from nhvr_b200 import ops
    from oracle.texture import texture_sample, composite
    torch.manual_seed(71)
    uvp = (torch.randn(2, 73, 20, 24, device=cuda_dev)).requires_grad_(True)
    atlas = smooth_atlas(3, 16, cuda_dev).requires_grad_(True)
    u2, a2 = uvp.detach().clone().requires_grad_(True), atlas.detach().clone().requires_grad_(True)
    w = torch.randn(2, 3, 20, 24, device=cuda_dev)
    for use_mask in (True, False):        # --use_mask_texture (start.sh) and the renormalised blend pretrain_start.sh trains with
        for t in (uvp, atlas, u2, a2):
            t.grad = None
        (ops.texture_sample_diff(uvp, atlas, use_mask) * w).sum().backward()
        (texture_sample(u2, a2, use_mask)[0] * w).sum().backward()
        assert (uvp.grad - u2.grad).abs().max().item() <= 1e-3 * u2.grad.abs().max().item(), use_mask
        assert (atlas.grad - a2.grad).abs().max().item() <= 1e-3 * a2.grad.abs().max().item(), use_mask
    fgm = torch.rand(3, 4, 10, 12, device=cuda_dev).requires_grad_(True)
    bg = torch.rand(3, 10, 12, device=cuda_dev).requires_grad_(True)
    f2, b2 = fgm.detach().clone().requires_grad_(True), bg.detach().clone().requires_grad_(True)
    w = torch.randn(3, 3, 10, 12, device=cuda_dev)
    (ops.composite_diff(fgm, bg) * w).sum().backward()
    (composite(f2, b2) * w).sum().backward()
    assert (fgm.grad - f2.grad).abs().max().item() <= 1e-5 and (bg.grad - b2.grad).abs().max().item() <= 1e-5


def test_sampler_backward_flat_uv_takes_the_aggregated_path(cuda_dev):
    """Stick-figure pose maps give IUV that is constant over most of a frame: whole warps then add to the same four texels of
    every part and nhvr_texture_sample_bwd sums a warp's contributions before ONE lane issues the reductions.  Flat, half-flat
    (warps that mix uniform and varying lanes take the per-lane path) and ragged-tail shapes against autograd on the oracle."""
    from nhvr_b200 import ops
    from oracle.texture import texture_sample
    torch.manual_seed(72)
    for (N, H, W) in ((2, 16, 64), (1, 9, 50)):               # W = 50: warps straddle rows and the last warp is partial
        base = torch.randn(1, 73, 1, 1, device=cuda_dev).expand(N, 73, H, W).clone()
        base[:, :, H // 2:, : W // 2] = torch.randn(N, 73, H - H // 2, W // 2, device=cuda_dev)      # one varying quadrant
        base[0, 25:, 0, :32] = 5.0                                                                     # a full warp at u = v = 1 (clamped)
        uvp = base.clone().requires_grad_(True)
        atlas = smooth_atlas(3, 16, cuda_dev, seed=5).requires_grad_(True)
        u2, a2 = uvp.detach().clone().requires_grad_(True), atlas.detach().clone().requires_grad_(True)
        w = torch.randn(N, 3, H, W, device=cuda_dev)
        for use_mask in (True, False):
            for t in (uvp, atlas, u2, a2):
                t.grad = None
            (ops.texture_sample_diff(uvp, atlas, use_mask) * w).sum().backward()
            (texture_sample(u2, a2, use_mask)[0] * w).sum().backward()
            assert (uvp.grad - u2.grad).abs().max().item() <= 1e-3 * u2.grad.abs().max().item(), (N, H, W, use_mask)
            assert (atlas.grad - a2.grad).abs().max().item() <= 1e-3 * a2.grad.abs().max().item(), (N, H, W, use_mask)


def test_discriminator_backward_parity(cuda_dev):
    """D and G objectives through the multiscale PatchGAN (LSGAN + feature matching): gradients w.r.t. the
    discriminator's parameters and w.r.t. its input image (what the generator receives) vs torch autograd."""
    from nhvr_b200.networks import define_D
    from nhvr_b200 import losses as L
    from oracle.networks import define_D as oracle_define_D
    from oracle import losses as O
    torch.manual_seed(81)
    ref = oracle_define_D(6, 16, 3, "instance", False, 2, True).to(cuda_dev)
    net = define_D(6, 16, 3, "instance", False, 2, True)
    net.load_state_dict(ref.state_dict())
    fake = (torch.rand(2, 6, 66, 66, device=cuda_dev) * 2 - 1).requires_grad_(True)
    real = torch.rand(2, 6, 66, 66, device=cuda_dev) * 2 - 1
    fake_r = fake.detach().clone().requires_grad_(True)
    with torch.no_grad():
        pr = net(real)
        pr_r = ref(real)
    pf = net(fake)
    loss = L.lsgan_diff(pf, True) + L.feature_matching_diff(pf, pr, 3, 2, 10.0)
    loss.backward()
    pf_r = ref(fake_r)
    loss_r = O.gan_loss(pf_r, True) + O.feature_matching_loss(pf_r, pr_r, 3, 2, 10.0)
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) <= 2e-2 * abs(loss_r.item())
    rows = [("input", fake.grad, fake_r.grad)] + [(k, p.grad, q.grad) for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters())]
    for name, a, b in rows:
        assert a is not None, name
        if name.endswith(".bias") and a.abs().max().item() == 0.0:
            continue
        cos = torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
        assert cos >= 0.99, (name, cos)


def test_forward_without_backward_does_not_pin_engines(cuda_dev):
    """Grad-enabled forwards that are never back-propagated release their training engine when autograd frees the node."""
    net, _ = _pair_G(cuda_dev, 3, 3, 16, "global", 1, 1, seed=43)
    x = torch.rand(1, 3, 32, 32, device=cuda_dev) * 2 - 1
    for _ in range(12):                      # more than the 8-engine pool
        y = net(x)
        assert y.requires_grad
        del y
    pool = next(v for k, v in net._engines.items() if k[-1] == "train")
    assert len(pool) <= 2


@pytest.mark.parametrize("poses", ["smooth", "stick_figures"])
def test_full_training_step_runs_and_matches_oracle_losses(cuda_dev, poses):
    """configs[2] in miniature: one RenderTrainer step (D step + G step) — its loss values against the oracle's
    formulas on the same weights/batch, and a few steps of Adam change every parameter group.  `stick_figures`: the pose maps are
    the reference's real input (BODY_25 stick figures rasterised from the bundled keypoints, ~98 % flat background): InstanceNorm of
    nearly constant planes in the fp16 training engines, and the warp-aggregated atlas reductions of the lookup's backward."""
    from nhvr_b200.networks import define_D
    from nhvr_b200.train import RenderTrainer, synthetic_train_batch
    from oracle import losses as O
    from oracle.networks import define_D as oracle_define_D
    from oracle.texture import texture_sample, composite
    pipe, ref = _pair_pipeline(cuda_dev, _small_kw(), seed=91)
    torch.manual_seed(92)
    refD = oracle_define_D(6, 16, 3, "instance", False, 2, True).to(cuda_dev)
    netD = define_D(6, 16, 3, "instance", False, 2, True)
    netD.load_state_dict(refD.state_dict())
    batch = synthetic_train_batch(2, 64, cuda_dev, seed=5)
    if poses == "stick_figures":
        import numpy as np
        from nhvr_b200 import pose as posemod
        kps = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keypoints_body25.npy"))
        maps = torch.from_numpy(posemod.pose_maps(kps[[0, 1, 40, 41]], 64, 3)).to(cuda_dev)        # frames (t-1, t) of two samples
        batch["pose_prev"], batch["pose"] = maps[[0, 2]].contiguous(), maps[[1, 3]].contiguous()
    # oracle losses at the initial weights
    with torch.no_grad():
        def frame(pose, prev):
            uvp = ref.netTransG(pose)
            tex, _, _ = texture_sample(uvp, ref.atlas, True)
            fgm = ref.netG(torch.cat([tex, pose, prev], 1))
            return uvp, composite(fgm, ref.netBG(ref.bg.unsqueeze(0))[0])
        _, out0 = frame(batch["pose_prev"], torch.zeros_like(batch["image"]))
        uvp1, out1 = frame(batch["pose"], out0)
        pf, pr = refD(torch.cat([batch["pose"], out1], 1)), refD(torch.cat([batch["pose"], batch["image"]], 1))
        lD = 0.5 * (O.gan_loss(pf, False) + O.gan_loss(pr, True))
        # generator objective with the SAME discriminator weights (pix2pixHD computes both sides before either update)
        lG = (O.gan_loss(pf, True) + O.feature_matching_loss(pf, pr, 3, 2, 10.0) + 500.0 * O.l2_loss(out1, batch["image"])
              + 1000.0 * O.uv_loss(uvp1, batch["dp_i"], batch["dp_uv"]) + 10.0 * O.prob_loss(uvp1, batch["dp_i"])
              + 500.0 * O.temporal_loss(out1, out0, batch["flow_inv"]))
    trainer = RenderTrainer(pipe, netD)
    before = {k: v.detach().clone() for k, v in list(pipe.named_parameters()) + list(netD.named_parameters())}
    out = trainer.step(batch)
    assert abs(out["loss_D"].item() - lD.item()) <= 2e-2 * abs(lD.item()), (out["loss_D"].item(), lD.item())
    assert abs(out["loss_G"].item() - lG.item()) <= 2e-2 * abs(lG.item()), (out["loss_G"].item(), lG.item())
    for _ in range(2):
        out = trainer.step(batch)
    assert torch.isfinite(out["loss_G"]) and torch.isfinite(out["loss_D"])
    from nhvr_b200 import capi
    capi.check_overflow()                 # no fp16 overflow of a conv output / gradient on the way
    changed = [k for k, v in list(pipe.named_parameters()) + list(netD.named_parameters())
               if not k.endswith(".bias") and (v.detach() - before[k]).abs().max().item() > 0]
    assert any(k.startswith("netTransG") for k in changed) and any(k.startswith("netG") for k in changed)
    assert any(k.startswith("netBG") for k in changed) and "atlas" in changed and "bg" in changed
    assert any(k.startswith("scale0") for k in changed) and any(k.startswith("scale1") for k in changed)


def test_pose_rasteriser_bit_exact_on_bundled_keypoints(cuda_dev):
    """GPU keypoint -> pose-map rasteriser against the host module on all 100 bundled OpenPose frames (configs[0] input),
    512^2 with the 6 pose channels of start.sh and 256^2 with 3: bit-identical maps."""
    import numpy as np
    from nhvr_b200 import ops, pose as posemod
    kps = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keypoints_body25.npy"))
    for size, nc, sl in ((512, 6, slice(0, 100, 7)), (256, 3, slice(0, 100, 9)), (1024, 6, slice(3, 5))):
        host = torch.from_numpy(posemod.pose_maps(kps[sl], size, nc))
        dev = ops.pose_rasterize(torch.from_numpy(kps[sl]).to(cuda_dev), size, nc)
        assert torch.equal(dev.cpu(), host), (size, (dev.cpu() != host).sum().item())
    # a missing joint (confidence 0) drops its limbs on both sides
    k2 = kps[:2].copy(); k2[:, 4, 2] = 0.0; k2[:, 0] = 0.0
    assert torch.equal(ops.pose_rasterize(torch.from_numpy(k2).to(cuda_dev), 512, 3).cpu(), torch.from_numpy(posemod.pose_maps(k2, 512, 3)))


def test_unfold_texture_matches_oracle_and_inverts_the_lookup(cuda_dev):
    """unfold_texture (README.md:64) as the lookup's adjoint: GPU scatter against the fp64 oracle; and a round trip -
    frames rendered by looking a known atlas up are unfolded back to it wherever the UVs cover a texel."""
    from nhvr_b200 import ops
    from oracle.texture import unfold_texture
    torch.manual_seed(17)
    N, H, W, S = 3, 96, 80, 24
    img = torch.rand(N, 3, H, W) * 2 - 1
    dp_i = torch.randint(0, 25, (N, H, W))
    dp_uv = torch.rand(N, 2, H, W)
    unf = ops.TextureUnfolder(S, 3, cuda_dev)
    unf.add(img[:2].to(cuda_dev), dp_i[:2].to(cuda_dev), dp_uv[:2].to(cuda_dev))
    unf.add(img[2:].to(cuda_dev), dp_i[2:].to(cuda_dev), dp_uv[2:].to(cuda_dev))          # accumulates over batches
    got = unf.atlas().cpu()
    ref = unfold_texture(img, dp_i, dp_uv, S)
    assert got.shape == ref.shape == (24, 3, S, S)
    assert (got - ref).abs().max().item() <= 1e-4
    # round trip with texel-centred UVs: the weighted mean returns the atlas exactly where covered
    atlas = torch.rand(24, 3, S, S) * 2 - 1
    ys, xs = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    uv = torch.stack([xs.float() / (S - 1), ys.float() / (S - 1)], 0)                        # [2,S,S]
    parts = torch.arange(1, 25).view(24, 1, 1).expand(24, S, S)
    frames = atlas.clone()                                                                   # frame k shows part k's texture
    unf2 = ops.TextureUnfolder(S, 3, cuda_dev)
    unf2.add(frames.to(cuda_dev), parts.contiguous().to(cuda_dev), uv.unsqueeze(0).expand(24, 2, S, S).contiguous().to(cuda_dev))
    back = unf2.atlas().cpu()
    assert (back - atlas).abs().max().item() <= 1e-4


def test_native_adam_matches_torch(cuda_dev):
    """nhvr_adam_step over a flat bucket against torch.optim.Adam on the same gradients, 4 steps; and the engines see the
    update (packed weights are re-packed through the bucket's ext_version)."""
    from nhvr_b200.train import ParamBucket
    net, ref = _pair_G(cuda_dev, 3, 3, 16, "global", 1, 1, seed=71)
    ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=2e-4, betas=(0.5, 0.999))
    bucket = ParamBucket(net.parameters(), 2e-4, 0.5, owners=[net])
    x = torch.rand(2, 3, 32, 32, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        y0 = net(x)
    for _ in range(4):
        opt.zero_grad()
        ref(x).square().mean().backward()
        bucket.zero_grad()
        for p, q in zip(net.parameters(), ref.parameters()):
            p.grad.copy_(q.grad)                                  # identical gradients: isolates the optimiser
        opt.step()
        bucket.adam_step()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p, q, atol=2e-7, rtol=1e-5), k
    with torch.no_grad():
        y1, y1_ref = net(x), ref(x)
    assert (y1 - y0).abs().max().item() > 1e-5                     # the inference engine picked the new weights up
    assert (y1 - y1_ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("N,C,H,W", [(2, 4, 37, 53), (1, 73, 40, 24), (3, 1, 66, 66), (8, 4, 512, 512)])
def test_bias_grad_chunked(cuda_dev, N, C, H, W):
    """nhvr_bias_grad over (channel, chunk) blocks: odd plane sizes (chunk starts that are not 16-byte aligned take the scalar
    path), a single channel, and the RGB + mask head's shape."""
    from nhvr_b200 import ops
    g = torch.Generator().manual_seed(N * 100 + C)
    x = torch.randn(N, C, H, W, generator=g).to(cuda_dev)
    db = ops.bias_grad(x, 0.5)
    ref = x.double().sum((0, 2, 3)) * 0.5
    assert torch.allclose(db.double(), ref, rtol=1e-5, atol=1e-3 * (N * H * W) ** 0.5 * 1e-2)


def test_vgg19_features_and_perceptual_loss(cuda_dev):
    """pix2pixHD's VGGLoss (on unless --no_vgg_loss): the five VGG19 feature taps, the weighted L1 loss and its gradient with
    respect to the generated image, random weights on both sides (there is no ImageNet checkpoint offline)."""
    from nhvr_b200 import capi, losses as L
    from nhvr_b200.networks import Vgg19B200
    from oracle.networks import Vgg19
    from oracle.losses import vgg_loss
    prev = capi.operand_dtype()
    capi.set_operand_dtype("f16")
    try:
        torch.manual_seed(5)
        ref = Vgg19().to(cuda_dev).eval()
        net = Vgg19B200().to(cuda_dev)
        net.load_state_dict(ref.state_dict())
        x = (torch.rand(2, 3, 64, 96, device=cuda_dev) * 2 - 1).requires_grad_(True)
        y = torch.rand(2, 3, 64, 96, device=cuda_dev) * 2 - 1
        with torch.no_grad():
            fa, fb = net(x.detach()), ref(x.detach())
        for a, b in zip(fa, fb):
            assert a.shape == b.shape
            assert (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item()), (a.shape, (a - b).abs().max().item(), b.abs().max().item())
        loss = L.vgg_diff(net, x, y, 10.0)
        loss.backward()
        xr = x.detach().clone().requires_grad_(True)
        loss_r = 10.0 * vgg_loss(ref, xr, y)
        loss_r.backward()
        assert abs(loss.item() - loss_r.item()) <= 1e-2 * abs(loss_r.item()), (loss.item(), loss_r.item())
        cos = torch.nn.functional.cosine_similarity(x.grad.flatten().double(), xr.grad.flatten().double(), dim=0).item()
        rel = ((x.grad - xr.grad).double().norm() / xr.grad.double().norm()).item()
        assert cos >= 0.99 and rel <= 0.15, (cos, rel)
        # odd sizes: floor pooling
        with torch.no_grad():
            z = torch.rand(1, 3, 50, 38, device=cuda_dev)
            for a, b in zip(net(z), ref(z)):
                assert a.shape == b.shape and (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item())
    finally:
        capi.set_operand_dtype(prev)
