"""GPU parity of the split-precision ("3 x fp16") path and of the whole path FREE-RUNNING at the reference's real
configuration (test_start/start.sh: 512 x 512, ngf_global 48 / 2 down / 10 blocks, UV generator ngf 64 / 5 blocks,
--use_laplace --pose_plus_laplace => 6 pose channels, 200 x 200 atlas, the 100 bundled keypoint JSONs).

Bars (BASELINE.json north_star): rendered frames max-abs <= 2e-2 on [-1, 1] and PSNR >= 45 dB against the fp32 oracle,
free-running (the path's own UV-generator output feeds the lookup, its own frames feed back); integer part / texel
bit-exact on identical UV inputs (test_gpu_parity.py) and reported as agreement here.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def psnr(a, b, peak=2.0):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * math.log10(peak * peak / mse)


def smooth_atlas(C, S, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(24, C, 6, 6, generator=g)
    return torch.tanh(torch.nn.functional.interpolate(low, size=(S, S), mode="bicubic", align_corners=False)).to(dev)


@pytest.fixture(autouse=True)
def _f16(cuda_dev):
    from nhvr_b200 import capi
    prev = capi.operand_dtype()
    capi.set_operand_dtype("f16")
    yield
    capi.set_operand_dtype(prev)


# ------------------------------------------------------------------ hilo format
@pytest.mark.parametrize("pad,split,halo", [((3, 3, 3, 3), 0, "R"), ((1, 1, 1, 1), 1, "Z"), ((0, 0, 1, 1), 0, "Z")])
def test_hilo_pack_unpack(cuda_dev, pad, split, halo):
    """fp32 -> (hi, lo) fp16 planes -> fp32: 22 significant bits (2^-22 relative, 2^-25 absolute in the subnormal range)."""
    from nhvr_b200 import ops, capi
    torch.manual_seed(9)
    x = torch.randn(2, 21, 13, 17, device=cuda_dev) * torch.logspace(-3, 2, 21, device=cuda_dev).view(1, 21, 1, 1)
    buf = ops.P8Buffer(ops.make_desc(2, 4, 13, 17, pad, split, capi.HALO_REFLECT if halo == "R" else capi.HALO_ZERO, hilo=1))
    assert buf.desc.C8 == 8
    ops.pack_nchw([x[:, :5], x[:, 5:]], buf)
    y = ops.unpack_nchw(buf, 21)
    err = (y - x).abs()
    assert (err <= x.abs() * 2.0 ** -21 + 2.0 ** -24).all(), (err / x.abs().clamp(min=1e-6)).max().item()


def _conv3_case(dev, kind, cin, cout, k, stride, pad, N, H, W, halo, epi, act=0, no_wlo=False):
    """One split-precision conv through the C-ABI against torch in fp64 on the SAME fp32 operands
    (no_wlo: conv flag bit 6 - the reference then takes the weights rounded to 16 bits, which is all that variant drops)."""
    import torch.nn.functional as F
    from nhvr_b200 import ops, capi
    plan = ops.ConvPlan(kind, cin, cout, k, stride, pad, N, H, W, halo, epi, act, split3=True, no_wlo=no_wlo)
    assert plan.in_desc.hilo == 1
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + k)
    x = (torch.rand(N, cin, H, W, generator=g) * 2 - 1).to(dev)
    transposed = kind == capi.CONV_TRANSPOSE
    w = (torch.randn(*((cin, cout, k, k) if transposed else (cout, cin, k, k)), generator=g) * (1.0 / (cin * k * k) ** 0.5)).to(dev)
    b = (torch.rand(cout, generator=g) - 0.5).to(dev)
    xin = ops.P8Buffer(plan.in_desc.copy(), dev)
    ops.pack_nchw([x], xin)
    plan.pack_weights(w)
    xd, wd = x.double(), (w.half() if no_wlo else w).double()
    if transposed:
        ref = F.conv_transpose2d(xd, wd, stride=2, padding=pad, output_padding=1 if k == 3 else 0)
    else:
        xp = F.pad(xd, (pad,) * 4, mode="reflect" if halo == capi.HALO_REFLECT else "constant")
        ref = F.conv2d(xp, wd, stride=stride)
    assert (plan.Ho, plan.Wo) == tuple(ref.shape[-2:])
    stat_err = 0.0
    if epi == capi.EPI_RAW_STATS:
        raw = ops.P8Buffer(plan.raw_desc(), dev)
        assert raw.desc.hilo == 1
        stats = torch.zeros(N * plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
        plan.forward(xin, raw.ptr, stats=stats)
        out = ops.unpack_nchw(raw, cout)
        st = stats.view(N, plan.Cout8 * 8, 4)[:, :cout, :2].double()
        s_ref = torch.stack([ref.sum((2, 3)), (ref ** 2).sum((2, 3))], -1)
        stat_err = ((st - s_ref).abs() / (1.0 + s_ref.abs())).max().item()
    else:
        out = torch.empty(N, cout, plan.Ho, plan.Wo, dtype=torch.float32, device=dev)
        plan.forward(xin, out.data_ptr(), bias=b)
        ref = ref + b.double().view(1, -1, 1, 1)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    return err, stat_err, plan.info()


@pytest.mark.parametrize("name,kind,cin,cout,k,stride,pad,N,H,W,halo,epi", [
    ("res_256", "CONV", 256, 256, 3, 1, 1, 1, 40, 44, "R", "RAW_STATS"),            # 16 chunks, two slab stages
    ("res_64_small", "CONV", 64, 64, 3, 1, 1, 2, 19, 23, "R", "RAW_STATS"),
    ("down_s2", "CONV", 64, 128, 3, 2, 1, 2, 130, 128, "Z", "RAW_STATS"),           # parity-split hilo input, single slab stage
    ("down_s2_256", "CONV", 128, 256, 3, 2, 1, 1, 66, 70, "Z", "RAW_STATS"),
    ("up_convT", "CONV_TRANSPOSE", 256, 128, 3, 2, 1, 1, 24, 40, "Z", "RAW_STATS"), # 4 phase accumulators, N split
    ("up_convT_64", "CONV_TRANSPOSE", 128, 64, 3, 2, 1, 2, 33, 31, "Z", "RAW_STATS"),
    ("stem_6ch", "CONV", 6, 64, 7, 1, 3, 2, 224, 128, "R", "RAW_STATS"),            # one logical plane + zero plane, M-replicated
    ("head_73", "CONV", 64, 73, 7, 1, 3, 2, 96, 128, "R", "BIAS_ACT_F32"),          # single slab stage x 4 chunks, fp32 NCHW output
    ("head_73_odd", "CONV", 64, 73, 7, 1, 3, 1, 45, 150, "R", "BIAS_ACT_F32"),
    ("head_rgb_rowmode", "CONV", 48, 4, 7, 1, 3, 2, 130, 300, "R", "BIAS_ACT_F32"),   # row mode + stacked rows + split precision
])
def test_conv_split3(cuda_dev, name, kind, cin, cout, k, stride, pad, N, H, W, halo, epi):
    """hi/lo operands, 3 MMAs per K step: the result is an fp32-accumulated product of (nearly) un-rounded operands -
    within 2e-5 of the fp64 result relative to its scale (a single fp16 operand pair gives ~5e-4)."""
    from nhvr_b200 import capi
    err, stat_err, info = _conv3_case(cuda_dev, getattr(capi, kind), cin, cout, k, stride, pad, N, H, W,
                                      capi.HALO_REFLECT if halo == "R" else capi.HALO_ZERO, getattr(capi, "EPI_" + epi))
    assert info["kcp"] % 4 == 0, info
    assert err <= 2e-5, (name, err, info)
    assert stat_err <= 5e-4, (name, stat_err)       # fp32 atomics over H*W values


@pytest.mark.parametrize("name,kind,cin,cout,k,stride,pad,N,H,W,halo,epi", [
    ("res_192", "CONV", 192, 192, 3, 1, 1, 1, 40, 44, "R", "RAW_STATS"),
    ("down_s2", "CONV", 48, 96, 3, 2, 1, 2, 130, 128, "Z", "RAW_STATS"),
    ("up_convT", "CONV_TRANSPOSE", 192, 96, 3, 2, 1, 1, 24, 40, "Z", "RAW_STATS"),
    ("stem_12ch", "CONV", 12, 48, 7, 1, 3, 2, 224, 128, "R", "RAW_STATS"),
    ("stem_6ch_khalf", "CONV", 6, 64, 7, 1, 3, 1, 96, 128, "R", "RAW_STATS"),
    ("head_rgb_rowmode", "CONV", 48, 4, 7, 1, 3, 2, 130, 300, "R", "BIAS_ACT_F32"),
    ("bg_head", "CONV", 48, 3, 7, 1, 3, 1, 64, 64, "R", "BIAS_ACT_F32"),
])
def test_conv_split2(cuda_dev, name, kind, cin, cout, k, stride, pad, N, H, W, halo, epi):
    """Conv flag bit 6: hi/lo activations, weights rounded to 16 bits, 2 MMAs per K step (the temporal generator of the
    "strict2" preset).  Against fp64 on the fp32 activations and the 16-bit-rounded weights: the same 2e-5 as the 3-MMA form."""
    from nhvr_b200 import capi
    err, stat_err, info = _conv3_case(cuda_dev, getattr(capi, kind), cin, cout, k, stride, pad, N, H, W,
                                      capi.HALO_REFLECT if halo == "R" else capi.HALO_ZERO, getattr(capi, "EPI_" + epi), no_wlo=True)
    assert err <= 2e-5, (name, err, info)
    assert stat_err <= 5e-4, (name, stat_err)


def test_in_apply_hilo(cuda_dev):
    """InstanceNorm + ReLU + residual + mirrored halo on hilo activations against torch in fp64."""
    from nhvr_b200 import ops, capi
    torch.manual_seed(4)
    N, Cc, H, W = 2, 32, 21, 37
    raw_t = torch.randn(N, Cc, H, W, device=cuda_dev) * 3 + 0.5
    res_t = torch.randn(N, Cc, H, W, device=cuda_dev)
    raw = ops.P8Buffer(ops.make_desc(N, 4, H, W, hilo=1))
    res = ops.P8Buffer(ops.make_desc(N, 4, H, W, (1, 1, 1, 1), 0, capi.HALO_REFLECT, hilo=1))
    ops.pack_nchw([raw_t], raw)
    ops.pack_nchw([res_t], res)
    z = torch.zeros(N, Cc, dtype=torch.float64, device=cuda_dev)
    stats = torch.stack([raw_t.double().sum((2, 3)), (raw_t.double() ** 2).sum((2, 3)), z, z], -1).contiguous().view(-1)
    for act, with_res, pad, split, halo in [(capi.ACT_RELU, False, (1, 1, 1, 1), 0, capi.HALO_REFLECT),
                                            (capi.ACT_NONE, True, (1, 1, 1, 1), 0, capi.HALO_REFLECT),
                                            (capi.ACT_RELU, False, (1, 1, 1, 1), 1, capi.HALO_ZERO),
                                            (capi.ACT_RELU, False, (3, 3, 3, 3), 0, capi.HALO_REFLECT)]:
        dst = ops.P8Buffer(ops.make_desc(N, 4, H, W, pad, split, halo, hilo=1))
        ops.in_apply(raw, stats, act, dst, residual=res if with_res else None)
        y = ops.unpack_nchw(dst, Cc)
        ref = torch.nn.functional.instance_norm(raw_t.double(), eps=1e-5)
        if act == capi.ACT_RELU:
            ref = torch.relu(ref)
        if with_res:
            ref = ref + res_t.double()
        assert (y.double() - ref).abs().max().item() <= 2e-5, (act, with_res, (y.double() - ref).abs().max().item())


# ------------------------------------------------------------------ UV generator in split precision
@pytest.mark.parametrize("cin,ngf,nb,size,batch", [(3, 32, 3, 96, 1), (6, 64, 5, 256, 2)])
def test_uv_generator_split3_parity(cuda_dev, cin, ngf, nb, size, batch):
    """The 17-conv UV generator (second case: the reference's widths, pretrainTrans.sh) in split precision: raw 73-channel
    output within 2e-4 of the fp32 oracle's (the single-fp16 chain is at ~1.6e-2)."""
    from nhvr_b200.networks import define_G
    from oracle.networks import define_G as oracle_define_G
    torch.manual_seed(0)
    ref = oracle_define_G(cin, 73, ngf, "translate", 2, nb).to(cuda_dev).eval()
    net = define_G(cin, 73, ngf, "translate", 2, nb)
    net.load_state_dict(ref.state_dict())
    net.set_precision("split3")
    torch.manual_seed(1)
    x = torch.rand(batch, cin, size, size, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        y, y_ref = net(x), ref(x)
    err = (y - y_ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, y_ref.abs().max().item()), err
    net.set_precision("f16")
    with torch.no_grad():
        y16 = net(x)
    assert (y16 - y_ref).abs().max().item() > err            # the fast mode really is a different engine


# ------------------------------------------------------------------ the real configuration, free-running
START_SH_KW = dict(pose_nc=6, tex_nc=3, size=512, atlas_size=200, ngf_global=48, n_downsample_global=2, n_blocks_global=10,
                   ngf_translate=64, n_downsample_translate=2, n_blocks_translate=5, ngf_bg=48, n_downsample_bg=2, n_blocks_bg=2,
                   use_mask_texture=True)


def bundled_pose_maps(n_frames, size=512, pose_nc=6):
    """Pose maps of the reference's own keypoint fixtures (configs[0]), rasterised by the host-side pose module."""
    from nhvr_b200 import pose as posemod
    kps = np.load(os.path.join(GOLD, "keypoints_body25.npy"))[:n_frames]
    return torch.from_numpy(posemod.pose_maps(kps, size, pose_nc))


def _start_sh_pair(dev, atlas_kind, **prec):
    from nhvr_b200.pipeline import RenderPipeline
    from oracle.pipeline import RenderModel
    torch.manual_seed(11)
    ref = RenderModel(**START_SH_KW).to(dev).eval()
    if atlas_kind == "smooth":
        with torch.no_grad():
            ref.atlas.copy_(smooth_atlas(3, 200, dev))
    pipe = RenderPipeline(**START_SH_KW, **prec).to(dev)
    pipe.load_state_dict(ref.state_dict())
    return pipe, ref


def _oracle_frames(ref, poses):
    with torch.no_grad():
        bg_r = ref.refine_bg()
        prev = torch.zeros(1, 3, 512, 512, device=poses.device)
        outs = []
        for t in range(poses.shape[0]):
            o = ref.render_frame(poses[t:t + 1], prev, bg_r)
            prev = o["out"]
            outs.append(o)
    return outs


@pytest.mark.parametrize("precision,atlas_kind", [("strict", "smooth"), ("strict", "uniform"), ("strict2", "smooth"), ("balanced", "smooth")])
def test_frame_parity_start_sh_configuration(cuda_dev, precision, atlas_kind):
    """The frame function of `bash test_start/start.sh`'s configuration on the 8 first bundled keypoint frames: for every
    frame t the path renders (pose_t, previous frame of the ORACLE) - "the same inputs and weights" of north_star - with its
    own UV generator output feeding its own lookup and generator (nothing teacher-forced inside the frame).

    strict / texture-like atlas: north_star's bar, max-abs <= 2e-2 and PSNR >= 45 dB, on every frame (measured ~1.5e-3).
    strict / U(-1,1) white-noise atlas (the bench's): a UV error e moves the lookup by 100 e texels of independent noise, so
      the fp32 oracle's OWN rounding noise (against an fp64 evaluation of the same model, computed here) already costs
      4e-3 .. 7e-3 of the 2e-2; held to PSNR >= 45 dB, max-abs <= 6e-2 and <= 12 x that reference noise floor.
    strict2 (temporal generator with hi + lo activations but 16-bit weights, 2 MMAs per product) / texture-like atlas: the same
      north_star bar, 2e-2 / 45 dB, with a smaller margin (measured 5.5e-3 .. 7.9e-3, 74-75 dB) - additionally held to 1.2e-2.
    balanced (temporal generator in plain fp16) / texture-like atlas: PSNR >= 60 dB; max-abs <= 4e-2 (fp16 is a RELATIVE
      precision, and InstanceNorm of a 98 %-flat stick-figure map puts |z| ~ 25-50 on the limb pixels: the few pixels next
      to them sit at 1e-2 .. 2.5e-2, run-to-run)."""
    pipe, ref = _start_sh_pair(cuda_dev, atlas_kind, precision=precision)
    assert pipe.precision == precision
    T = 8
    poses = bundled_pose_maps(T).to(cuda_dev)
    refs = _oracle_frames(ref, poses)
    floor = None
    if atlas_kind == "uniform":
        from oracle.pipeline import RenderModel
        ref64 = RenderModel(**START_SH_KW).to(cuda_dev).double().eval()
        ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
        with torch.no_grad():
            bg64 = ref64.refine_bg()
            floor = []
            for t in range(T):
                prev = refs[t - 1]["out"].double() if t else torch.zeros(1, 3, 512, 512, device=cuda_dev, dtype=torch.float64)
                o64 = ref64.render_frame(poses[t:t + 1].double(), prev, bg64)
                floor.append((refs[t]["out"].double() - o64["out"]).abs().max().item())
        del ref64
    worst, worst_db, agree = 0.0, 99.0, 1.0
    with torch.no_grad():
        bg = pipe.refine_bg()
        for t in range(T):
            prev = refs[t - 1]["out"] if t else torch.zeros(1, 3, 512, 512, device=cuda_dev)
            r = pipe.render_frame(poses[t:t + 1], prev, bg)
            e = (r["out"] - refs[t]["out"]).abs().max().item()
            db = psnr(r["out"], refs[t]["out"])
            agree = min(agree, (r["part"] == refs[t]["part"]).float().mean().item())
            worst, worst_db = max(worst, e), min(worst_db, db)
            if precision == "balanced":
                assert e <= 4e-2 and db >= 60.0, (t, e, db)
            elif atlas_kind == "smooth":
                assert e <= (1.2e-2 if precision == "strict2" else 2e-2) and db >= 45.0, (t, e, db)
            else:
                assert db >= 45.0 and e <= 6e-2 and e <= 12.0 * floor[t], (t, e, db, floor[t])
    from nhvr_b200 import capi
    capi.check_overflow(cuda_dev)
    print("[frame parity, %s, %s atlas] worst max-abs %.3e, worst PSNR %.1f dB, part agreement %.6f%s"
          % (precision, atlas_kind, worst, worst_db, agree,
             "" if floor is None else ", fp32-oracle noise floor %.2e..%.2e" % (min(floor), max(floor))))
    assert agree >= 0.9999


def test_free_running_start_sh_configuration(cuda_dev):
    """Free-running (the path's own frames fed back) at the same configuration.  With random-init weights the reference's
    temporal recurrence is CHAOTIC: a 1e-6 perturbation of the first previous-frame input of the fp32 oracle itself grows
    ~2x per frame (measured here, `growth`), so any two implementations - including the oracle on two different
    libraries - separate exponentially and the 2e-2 bar can only hold over a horizon set by the first frame's error.
    Asserted (default "strict" precision, texture-like atlas, public clip API = CUDA-graph replay): (1) >= 4 free-running
    frames within 2e-2 / 45 dB; (2) the path adds no instability of its own: its error grows no faster than 1.5 x the
    oracle's own perturbation growth; (3) the "balanced" preset keeps its first two frames >= 45 dB."""
    T = 6
    poses = bundled_pose_maps(T).to(cuda_dev)
    pipe, ref = _start_sh_pair(cuda_dev, "smooth")
    refs = _oracle_frames(ref, poses)
    frames_ref = torch.cat([o["out"] for o in refs])
    with torch.no_grad():
        bg_r = ref.refine_bg()
        prev_p = 1e-6 * torch.randn(1, 3, 512, 512, device=cuda_dev)
        self_div = []
        for t in range(T):
            prev_p = ref.render_frame(poses[t:t + 1], prev_p, bg_r)["out"]
            self_div.append((prev_p - refs[t]["out"]).abs().max().item())
        frames = pipe.render_clip(poses)                       # public API, CUDA graph replay
    errs = [(frames[t] - frames_ref[t]).abs().max().item() for t in range(T)]
    dbs = [psnr(frames[t], frames_ref[t]) for t in range(T)]
    growth = (self_div[-1] / self_div[1]) ** (1.0 / (T - 2))
    own = (errs[-1] / errs[1]) ** (1.0 / (T - 2))
    print("[free-running] strict max-abs per frame %s | PSNR %s | oracle self-divergence %s | growth/frame: oracle %.2f, path %.2f"
          % (" ".join("%.1e" % e for e in errs), " ".join("%.0f" % d for d in dbs), " ".join("%.1e" % e for e in self_div), growth, own))
    for t in range(4):
        assert errs[t] <= 2e-2 and dbs[t] >= 45.0, (t, errs[t], dbs[t])
    assert own <= 1.5 * growth, (own, growth)
    pipe2, _ = _start_sh_pair(cuda_dev, "smooth", precision="balanced")
    f0 = pipe2.render_clip(poses[:2])
    assert psnr(f0[0], frames_ref[0]) >= 45.0 and psnr(f0[1], frames_ref[1]) >= 45.0


def test_lockstep_8_clips_512_match_oracle_per_clip(cuda_dev):
    """The engines bench.py times (8 lock-step clips at 512^2 as the batch dimension, CUDA graph, texture-like atlas): every
    clip equals the oracle run on that clip alone - first frame (identical inputs) within 2e-2 / 45 dB, second frame
    (free-running, see test_free_running_start_sh_configuration for the recurrence's own error growth) >= 45 dB."""
    from nhvr_b200 import pose as posemod
    from nhvr_b200.pipeline import RenderPipeline
    from oracle.pipeline import RenderModel
    kw = dict(START_SH_KW, pose_nc=3)
    torch.manual_seed(21)
    ref = RenderModel(**kw).to(cuda_dev).eval()
    with torch.no_grad():
        ref.atlas.copy_(smooth_atlas(3, 200, cuda_dev))
    pipe = RenderPipeline(**kw).to(cuda_dev)
    pipe.load_state_dict(ref.state_dict())
    B, T = 8, 2
    kps = np.load(os.path.join(GOLD, "keypoints_body25.npy"))
    poses = torch.stack([torch.from_numpy(posemod.pose_maps(kps[10 * b:10 * b + T], 512, 3)) for b in range(B)]).to(cuda_dev)
    frames = pipe.render_clips(poses)
    worst, worst_db, worst_db1 = 0.0, 99.0, 99.0
    for b in range(B):
        fr = ref.render_clip(poses[b])
        worst = max(worst, (frames[b, 0] - fr[0]).abs().max().item())
        worst_db = min(worst_db, psnr(frames[b, 0], fr[0]))
        worst_db1 = min(worst_db1, psnr(frames[b, 1], fr[1]))
    print("[8 x 512^2 lock-step] frame 0: worst clip max-abs %.3e, PSNR %.1f dB; frame 1: PSNR %.1f dB" % (worst, worst_db, worst_db1))
    assert worst <= 2e-2 and worst_db >= 45.0 and worst_db1 >= 45.0
