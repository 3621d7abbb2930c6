"""CPU suite: C-ABI library loads and exports every declared symbol; conv plan builder (host-only);
P8 geometry; clip sharding incl. a world_size-2 gloo run; pose rasteriser; checkpoint naming."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from nhvr_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "nhvr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nhvr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(capi.SYMBOLS.keys()), declared ^ set(capi.SYMBOLS.keys())
    assert lib.nhvr_version() >= 100
    assert lib.nhvr_strerror(-1).decode().startswith("device is not compute capability 10")


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_compute_entry_points_fail_loudly_without_sm100():
    from nhvr_b200 import capi
    lib = capi.load()
    assert lib.nhvr_arch_ok() == -1
    with pytest.raises(capi.NhvrError):
        capi.require_device()
    # a compute call off-device must return the ARCH error, never succeed silently (dummy non-null pointers)
    buf = (C.c_float * 64)()
    st = lib.nhvr_composite(C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 0, 1, 2, 2, C.cast(buf, C.c_void_p), None)
    assert st == -1


def _plan(kind, cin, cout, k, s, p, N, H, W, halo, epi):
    from nhvr_b200 import ops
    return ops.ConvPlan(kind, cin, cout, k, s, p, N, H, W, halo, epi)


def test_conv_plans_of_the_reference_networks():
    """Plan builder is host code: tile counts, job/run programs and budgets for every layer shape of
    G_main (ngf 48), TransG (ngf 64) and D (ndf 64) at 512^2."""
    from nhvr_b200 import capi
    R, Z = capi.HALO_REFLECT, capi.HALO_ZERO
    cases = [
        (capi.CONV, 9, 48, 7, 1, 3, 1, 512, 512, R, 49, 1),
        (capi.CONV, 48, 96, 3, 2, 1, 1, 512, 512, Z, 9, 1),
        (capi.CONV, 96, 192, 3, 2, 1, 1, 256, 256, Z, 9, 1),
        (capi.CONV, 192, 192, 3, 1, 1, 1, 128, 128, R, 9, 1),
        (capi.CONV_TRANSPOSE, 192, 96, 3, 2, 1, 1, 128, 128, Z, 9, 4),
        (capi.CONV_TRANSPOSE, 96, 48, 3, 2, 1, 1, 256, 256, Z, 9, 4),
        (capi.CONV, 48, 4, 7, 1, 3, 1, 512, 512, R, 7, 1),     # row mode: one job per filter row, N = kw*8
        (capi.CONV, 64, 73, 7, 1, 3, 1, 512, 512, R, 49, 1),
        (capi.CONV, 256, 256, 3, 1, 1, 8, 128, 128, R, 9, 1),
        (capi.CONV, 6, 64, 4, 2, 2, 1, 512, 512, Z, 16, 1),
        (capi.CONV, 256, 512, 4, 1, 2, 1, 65, 65, Z, 16, 1),
        (capi.CONV, 512, 1, 4, 1, 2, 1, 66, 66, Z, 16, 1),
    ]
    for kind, cin, cout, k, s, p, N, H, W, halo, njobs, nacc in cases:
        pl = _plan(kind, cin, cout, k, s, p, N, H, W, halo, capi.EPI_RAW_STATS)
        info = pl.info()
        assert info["njobs"] == njobs and info["nacc"] == nacc
        assert info["smem_bytes"] <= 227 * 1024 and info["tmem_cols"] <= 512 and info["nacc"] * info["Npad"] <= info["tmem_cols"]
        assert info["Npad"] % 16 == 0 and info["Npad"] * info["nsplit"] >= cout
        assert info["kcp"] % 2 == 0 and info["kcp"] * info["nchunks"] * 8 >= cin
        assert info["nblocks"] == info["nchunks"] * info["njobs"] * info["kcp"] // 2
        if kind == capi.CONV_TRANSPOSE:
            assert (pl.Ho, pl.Wo) == (2 * H, 2 * W)
        else:
            assert (pl.Ho, pl.Wo) == ((H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1)
        # linearised tiling covers every valid output position with 128-row tiles
        assert info["tiles_per_img"] * 128 >= (pl.Ho if kind == capi.CONV else H) * 1
        assert pl.flops == pytest.approx(2.0 * k * k * cin * cout * (H * W if kind == capi.CONV_TRANSPOSE else pl.Ho * pl.Wo) * N)
    # the bottleneck conv: one merged run of 3 rows, 130 tiles per 128^2 image (~128 CTAs)
    pl = _plan(capi.CONV, 192, 192, 3, 1, 1, 1, 128, 128, R, capi.EPI_RAW_STATS)
    assert pl.info()["nruns"] == 1 and pl.info()["slab_units"] == 390 and pl.info()["tiles_per_img"] == 130
    assert pl.flops == pytest.approx(10.87e9, rel=1e-3)


def test_split_precision_plans_block_counts():
    """Split precision (conv flag bit 3) carries two weight blocks per group of four physical planes and tap (w_hi, w_lo), flag bit 6
    ("split2": no w_lo) one: the MMA table, the packed weight bytes and the presets that use them (host logic only)."""
    from nhvr_b200 import capi, ops
    from nhvr_b200.pipeline import PRECISION_PRESETS
    R = capi.HALO_REFLECT
    for (cin, cout, k, H) in ((192, 192, 3, 128), (256, 256, 3, 128), (48, 4, 7, 512), (64, 73, 7, 512)):
        p3 = ops.ConvPlan(capi.CONV, cin, cout, k, 1, k // 2, 1, H, H, R, capi.EPI_RAW_STATS if cout > 8 and cout != 73 else capi.EPI_BIAS_ACT_F32,
                          capi.ACT_NONE, split3=True)
        p2 = ops.ConvPlan(capi.CONV, cin, cout, k, 1, k // 2, 1, H, H, R, capi.EPI_RAW_STATS if cout > 8 and cout != 73 else capi.EPI_BIAS_ACT_F32,
                          capi.ACT_NONE, split3=True, no_wlo=True)
        i3, i2 = p3.info(), p2.info()
        assert p3.in_desc.hilo == 1 and p2.in_desc.hilo == 1
        assert i3["kcp"] % 4 == 0 and i2["kcp"] % 4 == 0
        assert i3["nblocks"] == i3["nchunks"] * i3["njobs"] * 2 * (i3["kcp"] // 4)
        assert i2["nblocks"] == i2["nchunks"] * i2["njobs"] * (i2["kcp"] // 4)
        assert i2["nchunks"] * i2["kcp"] == i3["nchunks"] * i3["kcp"]                   # the same physical planes
        assert p2.weight_bytes < p3.weight_bytes and p2.flops == p3.flops
    assert PRECISION_PRESETS["strict"] == ("split3", "split3") and PRECISION_PRESETS["strict2"] == ("split3", "split2")


def test_precision_presets_are_offered_consistently():
    """The preset names of RenderPipeline are the choices of test.py's --precision and of bench.py, and every non-default one has a
    parity note for the bench's `modes` block."""
    import bench
    from nhvr_b200.options import TestOptions
    from nhvr_b200.pipeline import PRECISION_PRESETS
    to = TestOptions()
    to.initialize()
    act = next(a for a in to.parser._actions if "--precision" in a.option_strings)
    assert sorted(act.choices) == sorted(PRECISION_PRESETS) and act.default == "strict"
    assert sorted(bench.PARITY_NOTE) == sorted(PRECISION_PRESETS)


def test_conv_plan_rejects_bad_shapes():
    from nhvr_b200 import capi
    with pytest.raises(capi.NhvrError):
        _plan(capi.CONV, 0, 8, 3, 1, 1, 1, 8, 8, 0, 0)
    with pytest.raises(capi.NhvrError):
        _plan(capi.CONV_TRANSPOSE, 8, 8, 4, 2, 1, 1, 8, 8, 0, 0)
    with pytest.raises(capi.NhvrError):
        _plan(capi.CONV, 8, 8, 9, 1, 4, 1, 8, 8, 0, 0)          # 81 taps > job table


def test_p8_geometry_bytes():
    from nhvr_b200 import ops, capi
    lib = capi.load()
    d = ops.make_desc(2, 3, 10, 12, (1, 1, 1, 1), 0, capi.HALO_REFLECT)
    assert lib.nhvr_act_bytes(C.byref(d)) == (2 * 3 * 12 * 14 + 2048) * 16
    d = ops.make_desc(1, 2, 9, 11, (1, 1, 1, 1), 1, capi.HALO_ZERO)       # split rounds 11x13 up to 12x14
    assert lib.nhvr_act_bytes(C.byref(d)) == (1 * 2 * 12 * 14 + 2048) * 16


def test_shard_frames_partitions_exactly():
    from nhvr_b200.pipeline import shard_frames
    for n, world, cpr in [(4096, 8, 1), (4096, 8, 4), (100, 8, 1), (100, 3, 2), (5, 8, 1), (0, 2, 1)]:
        seen = []
        for r in range(world):
            clips = shard_frames(n, world, r, cpr)
            assert len(clips) == cpr
            for a, b in clips:
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
        assert seen == list(range(n))                     # contiguous, ordered, no overlap, nothing lost
    assert shard_frames(4096, 8, 3, 1) == [(1536, 2048)]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from nhvr_b200.pipeline import shard_frames
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mine = shard_frames(101, world, rank, 2)
    # each rank "renders" its clips: frame value = index; max-over-ranks timing via all_reduce MAX as in bench.py
    frames = torch.zeros(101)
    for a, b in mine:
        frames[a:b] = torch.arange(a, b, dtype=torch.float32) + 1
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((frames.tolist(), t.item()))
    dist.destroy_process_group()


def test_clip_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames == [float(i + 1) for i in range(101)]      # every frame rendered by exactly one rank
    assert tmax == 2.0


def test_pose_rasteriser_on_reference_fixture():
    from nhvr_b200 import pose
    kp = pose.read_keypoints(os.path.join(GOLD, "keypoints_frame0.json"))
    assert kp.shape == (25, 3) and (kp[:, 2] > 0.4).all()          # all 25 joints detected
    assert 200 < kp[:, 0].min() and kp[:, 1].max() < 900
    m = pose.pose_maps(kp[None], 128, pose_nc=6)
    assert m.shape == (1, 6, 128, 128) and m.min() == -1.0 and m.max() <= 1.0
    assert (m[0, 3:] == 0).all()                                    # Laplace channels absent in the fixture
    drawn = (m[0, :3] > -1).any(0)
    assert 0.005 < drawn.mean() < 0.2
    ys, xs = np.nonzero(drawn)
    assert abs(xs.mean() - kp[:, 0].mean() / 8) < 12 and abs(ys.mean() - kp[:, 1].mean() / 8) < 20
    # alignment is the identity without a target and a pure scale/shift with one
    seq = np.stack([kp, kp])
    assert np.array_equal(pose.align_to_target(seq, None), seq)
    tgt = seq.copy(); tgt[:, :, :2] = tgt[:, :, :2] * 0.5 + 10
    al = pose.align_to_target(seq, tgt)
    assert np.allclose(al[:, :, :2], tgt[:, :, :2], atol=1e-2)


def test_checkpoint_naming_roundtrip(tmp_path):
    from nhvr_b200.pipeline import RenderPipeline
    from nhvr_b200.checkpoint import save_pipeline, load_pipeline, net_path
    kw = dict(size=16, atlas_size=4, ngf_global=8, n_blocks_global=1, ngf_translate=8, n_blocks_translate=1, ngf_bg=8, n_blocks_bg=1)
    a, b = RenderPipeline(**kw).cpu(), RenderPipeline(**kw).cpu()
    save_pipeline(a, str(tmp_path), 30)
    assert os.path.isfile(net_path(str(tmp_path), 30, "G")) and net_path("d", 30, "TransG").endswith("30_net_TransG.pth")
    assert load_pipeline(b, str(tmp_path), 30)
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v.cpu(), w.cpu()), k
    assert not load_pipeline(b, str(tmp_path), 31)


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from nhvr_b200.train import FlatGradBucket
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    params[2].grad = None                                     # a parameter that received no gradient on this rank
    b = FlatGradBucket(params)
    b.all_reduce_mean()
    if rank == 0:
        q.put([p.grad.tolist() for p in params])
    dist.destroy_process_group()


def test_flat_grad_bucket_allreduce_world_size_2_gloo():
    """The training path's only collective: one flat all-reduce averaging the gradients over ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g0, g1, g2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert g0 == [[1.5] * 4] * 3                               # mean of 1 and 2
    assert g1 == [0.0, 1.5, 3.0, 4.5, 6.0]
    assert g2 == [[0.0, 0.0], [0.0, 0.0]]


def test_train_and_test_scripts_build_checkpoint_compatible_pipelines():
    """train_start/pretrain_start.sh (--use_laplace, no --pose_plus_laplace, no --use_mask_texture) and
    test_start/start.sh (--use_laplace --pose_plus_laplace --use_mask_texture) must describe the same networks: a
    checkpoint written by train.py loads in test.py (state_dict keys and shapes identical)."""
    import json
    from nhvr_b200.options import TestOptions, TrainOptions, pipeline_kwargs
    from nhvr_b200.pipeline import RenderPipeline
    flags = json.load(open(os.path.join(GOLD, "ref_flags.json")))
    train_argv = next(c["argv"] for f, cmds in flags.items() for c in cmds if c["entry"] == "train.py")
    test_argv = next(c["argv"] for f, cmds in flags.items() for c in cmds if c["entry"] == "test.py")
    ot = TrainOptions().parse([a.replace("$DANCE", "d").replace("${DANCE}", "d") for a in train_argv])
    oe = TestOptions().parse([a.replace("$DANCE", "d").replace("${DANCE}", "d") for a in test_argv])
    assert ot.pose_nc == oe.pose_nc == 6
    assert not ot.use_mask_texture and oe.use_mask_texture
    kw_t, kw_e = pipeline_kwargs(ot), pipeline_kwargs(oe)
    small = dict(size=64, atlas_size=16, ngf_global=8, ngf_translate=8, ngf_bg=8, n_blocks_global=1, n_blocks_translate=1)
    kw_t.update(small); kw_e.update(small)
    if torch.cuda.is_available():
        pytest.skip("CPU-side construction check")
    pt, pe = RenderPipeline(**kw_t), RenderPipeline(**kw_e)
    pe.load_state_dict(pt.state_dict(), strict=True)


def test_graph_posenorm_cli_recovers_a_known_scale_and_translation(tmp_path):
    """data/data_prep/graph_posenorm.py with run_alignPose.sh's flags: a source sequence that is the target one scaled by
    0.8 about a point and shifted is aligned back onto it (ankles and body height within 2 px)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("graph_posenorm", os.path.join(ROOT, "data", "data_prep", "graph_posenorm.py"))
    gp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gp)
    kps = np.load(os.path.join(GOLD, "keypoints_body25.npy"))[:30]
    tdir, sdir, odir = tmp_path / "tgt", tmp_path / "src", tmp_path / "out"
    tdir.mkdir(); sdir.mkdir()
    base = json.load(open(os.path.join(GOLD, "keypoints_frame0.json")))
    src = kps.copy()
    src[:, :, 0] = (kps[:, :, 0] - 500.0) * 0.8 + 430.0
    src[:, :, 1] = (kps[:, :, 1] - 800.0) * 0.8 + 700.0
    for d, arr in ((tdir, kps), (sdir, src)):
        for i, k in enumerate(arr):
            j = json.loads(json.dumps(base))
            j["people"][0]["pose_keypoints_2d"] = [float(v) for v in k.reshape(-1)]
            json.dump(j, open(d / ("frame%05d_keypoints.json" % i), "w"))
    flags = json.load(open(os.path.join(GOLD, "ref_flags.json")))["data/data_prep/run_alignPose.sh"][0]["argv"]
    argv = list(flags)
    for name, val in (("--target_keypoints", str(tdir)), ("--source_keypoints", str(sdir)), ("--results", str(odir)), ("--source_frames", str(sdir))):
        argv[argv.index(name) + 1] = val
    argv[argv.index("--target_spread") + 1:argv.index("--target_spread") + 3] = ["700", "900"]
    argv[argv.index("--source_spread") + 1:argv.index("--source_spread") + 3] = ["600", "800"]
    out = gp.main(argv)
    assert out.shape == kps.shape and len(os.listdir(odir)) == 30
    ank_o, ank_t = 0.5 * (out[:, 11, 1] + out[:, 14, 1]), 0.5 * (kps[:, 11, 1] + kps[:, 14, 1])
    assert np.abs(np.median(ank_o) - np.median(ank_t)) <= 6.0
    h_o, h_t = ank_o - out[:, 0, 1], ank_t - kps[:, 0, 1]
    assert abs(np.median(h_o) / np.median(h_t) - 1.0) <= 0.02
    back = json.load(open(odir / "frame00000_keypoints.json"))
    assert len(back["people"][0]["pose_keypoints_2d"]) == 75


def test_texture_grid_layout_round_trip_and_tex_flags_parse():
    """unfold_texture.py's texture.jpg layout (4 x 6 parts) round-trips; pre_train_tex.sh's flags parse verbatim."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("unfold_texture", os.path.join(ROOT, "unfold_texture.py"))
    ut = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ut)
    atlas = (torch.randint(0, 256, (24, 3, 8, 8)).float() / 127.5 - 1.0)
    grid = ut.atlas_to_grid(atlas)
    assert grid.shape == (32, 48, 3)
    assert torch.allclose(ut.grid_to_atlas(grid, 8), atlas, atol=1e-6)
    assert (grid[8:16, 16:24] == ((atlas[8].permute(1, 2, 0) + 1) * 127.5).round().numpy().astype(np.uint8)).all()   # part 8 -> row 1, col 2
    from nhvr_b200.options import TrainOptions
    argv = json.load(open(os.path.join(GOLD, "ref_flags.json")))["pre_train_tex.sh"][0]["argv"]
    to = TrainOptions()
    to.initialize()
    to.parser.add_argument("--synthetic_steps", type=int, default=0)
    opt = to.parse(argv)
    assert opt.input_nc == 81 and opt.loadSize == 200 and opt.n_blocks_global == 5 and opt.ngf_global == 64 and opt.use_mask_texture


def test_param_bucket_matches_torch_adam_and_keeps_views():
    """train.ParamBucket (host arithmetic path): parameters and gradients live in flat buffers as views; three Adam steps
    equal torch.optim.Adam(lr 2e-4, betas (0.5, 0.999)); the 1 / world_size of a summed all-reduce is folded in."""
    from nhvr_b200.train import ParamBucket
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=2e-4, betas=(0.5, 0.999))
    bucket = ParamBucket(net.parameters(), 2e-4, 0.5)
    assert all(p.data_ptr() >= bucket.flat_p.data_ptr() for p in net.parameters()) and bucket.n % 4 == 0
    x = torch.randn(11, 7)
    for _ in range(3):
        bucket.zero_grad()
        net(x).square().mean().backward()                       # autograd accumulates INTO the flat views
        assert all(p.grad.data_ptr() >= bucket.flat_g.data_ptr() for p in net.parameters())
        bucket.adam_step()
        opt.zero_grad()
        ref(x).square().mean().backward()
        opt.step()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p, q, atol=1e-7, rtol=1e-5)


def _gloo_bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from nhvr_b200.train import ParamBucket
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    torch.manual_seed(0)                                         # identical initial weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 2))
    bucket = ParamBucket(net.parameters(), 2e-4, 0.5)
    g = torch.Generator().manual_seed(100 + rank)                # a different data shard per rank
    x = torch.randn(8, 6, generator=g)
    for _ in range(2):
        bucket.zero_grad()
        net(x).square().mean().backward()
        bucket.all_reduce_async()                                # what the networks' after_backward hook fires
        bucket.adam_step()                                       # waits, folds 1 / world into the update
    if rank == 0:
        q.put((bucket.flat_p.tolist(), x.tolist()))
    else:
        q.put(("other", bucket.flat_p.tolist(), x.tolist()))
    dist.destroy_process_group()


def test_param_bucket_data_parallel_world_size_2_gloo():
    """Two ranks with different batches: asynchronous all-reduce of the flat gradient bucket + Adam gives every rank the
    parameters of ONE process trained on the mean gradient of both shards."""
    import torch.multiprocessing as mp
    from nhvr_b200.train import ParamBucket
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120), q.get(timeout=120)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = next(g for g in got if g[0] != "other")
    r1 = next(g for g in got if g[0] == "other")
    assert r0[0] == r1[1]                                        # both ranks hold identical parameters
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 2))
    bucket = ParamBucket(net.parameters(), 2e-4, 0.5)
    xs = [torch.tensor(r0[1]), torch.tensor(r1[2])]
    for _ in range(2):
        bucket.zero_grad()
        (0.5 * (net(xs[0]).square().mean() + net(xs[1]).square().mean())).backward()
        bucket.adam_step()
    assert torch.allclose(bucket.flat_p, torch.tensor(r0[0]), atol=1e-7)
