"""conv + InstanceNorm + activation (+ residual) + halo in ONE kernel (nhvr_conv_forward_in_fused) against the two-kernel
path it replaces (nhvr_conv_forward RAW_STATS + nhvr_in_apply) and against torch in fp64.

The fused epilogue normalises the fp32 accumulators directly (the two-kernel path normalises the 16-bit / hilo raw tensor),
so it can only be closer to the fp64 result; both must agree within one operand rounding.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _f16(cuda_dev):
    from nhvr_b200 import capi
    prev = capi.operand_dtype()
    capi.set_operand_dtype("f16")
    yield
    capi.set_operand_dtype(prev)


def _full_values(buf):
    """Every unit of a (non-split) P8 buffer, halo included, as fp32 [N][logical planes][Hp*Wp][8] (hi + lo for hilo buffers)."""
    d = buf.desc
    Hp, Wp = d.H + d.pad_t + d.pad_b, d.W + d.pad_l + d.pad_r
    t = buf.mem[:d.N * d.C8 * Hp * Wp * 16].view(torch.float16).float().view(d.N, d.C8, Hp * Wp, 8)
    if d.hilo:
        t = t.view(d.N, d.C8 // 4, 4, Hp * Wp, 8)
        t = (t[:, :, 0:2] + t[:, :, 2:4]).reshape(d.N, d.C8 // 2, Hp * Wp, 8)
    return t


def _case(dev, C, N, H, W, act, with_res, split3, dst_pad, dst_halo):
    from nhvr_b200 import ops, capi
    plan = ops.ConvPlan(capi.CONV, C, C, 3, 1, 1, N, H, W, capi.HALO_REFLECT, capi.EPI_RAW_STATS, split3=split3)
    assert plan.in_fused_supported(), "fused epilogue not supported for this plan on this GPU"
    g = torch.Generator(device="cpu").manual_seed(C + H + (7 if with_res else 0))
    x = (torch.randn(N, C, H, W, generator=g)).to(dev)
    w = (torch.randn(C, C, 3, 3, generator=g) * (1.0 / (C * 9) ** 0.5)).to(dev)
    xin = ops.P8Buffer(plan.in_desc.copy(), dev)
    ops.pack_nchw([x], xin)
    plan.pack_weights(w)
    hilo = 1 if split3 else 0
    ddesc = ops.make_desc(N, plan.Cout8, H, W, dst_pad, 0, dst_halo, hilo=hilo)      # logical planes in, physical in the desc
    # two-kernel path
    raw = ops.P8Buffer(plan.raw_desc(), dev)
    st_a = torch.zeros(N * plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
    dst_a = ops.P8Buffer(ddesc.copy(), dev)
    plan.forward(xin, raw.ptr, stats=st_a)
    ops.in_apply(raw, st_a, act, dst_a, residual=xin if with_res else None)
    # fused path (twice)
    st_b = torch.zeros_like(st_a)
    dst_b = ops.P8Buffer(ddesc.copy(), dev)
    sync = torch.zeros(N, dtype=torch.int32, device=dev)
    for _ in range(2):
        st_b.zero_(); sync.zero_()
        plan.forward_in_fused(xin, st_b, act, dst_b, sync, residual=xin if with_res else None)
    torch.cuda.synchronize()
    # fp32 partial sums per CTA (atomic order varies), merged in fp64; the fused record holds two replicas of {sum, sum sq}
    ra, rb = st_a.view(-1, 4), st_b.view(-1, 4)
    assert torch.allclose(ra[:, :2], rb[:, :2] + rb[:, 2:], rtol=1e-5, atol=1e-4)
    # whole buffers, halo included, against each other
    a, b = _full_values(dst_a), _full_values(dst_b)
    ya, yb = ops.unpack_nchw(dst_a, C), ops.unpack_nchw(dst_b, C)
    # torch fp64 on the operands the kernel sees (fp32 values; split precision keeps ~22 bits, plain fp16 11)
    xd = x.double() if split3 else x.half().double()
    wd = w.double() if split3 else w.half().double()
    z = F.conv2d(F.pad(xd, (1, 1, 1, 1), mode="reflect"), wd)
    z = F.instance_norm(z, eps=1e-5)
    if act == capi.ACT_RELU:
        z = F.relu(z)
    if with_res:
        z = z + xd
    tol = 2e-5 if split3 else 4e-3
    err_a, err_b = (ya.double() - z).abs().max().item(), (yb.double() - z).abs().max().item()
    assert err_b <= tol * max(1.0, z.abs().max().item()), (err_a, err_b)
    assert err_b <= err_a * 1.5 + 1e-6, (err_a, err_b)
    # halo: every unit of the destination (mirrored rows / columns, corners) must match the two-kernel path's layout
    diff = (a - b).abs()
    lim = 1e-4 if split3 else 2e-2
    assert diff.max().item() <= lim * max(1.0, z.abs().max().item()), diff.max().item()
    # and the halo must be populated exactly where the two-kernel path populates it
    assert ((a != 0) ^ (b != 0)).float().mean().item() < 1e-3
    return err_a, err_b


@pytest.mark.parametrize("split3", [False, True])
@pytest.mark.parametrize("with_res,act", [(False, 1), (True, 0)])
def test_in_fused_resblock_layer(cuda_dev, split3, with_res, act):
    from nhvr_b200 import capi
    act = capi.ACT_RELU if act else capi.ACT_NONE
    _case(cuda_dev, 64, 2, 40, 56, act, with_res, split3, (1, 1, 1, 1), capi.HALO_REFLECT)


def test_in_fused_pair_width_and_zero_halo(cuda_dev):
    """192 channels = the CTA-pair lowering (cta_group::2); destination with a zero halo (the transposed conv's input)."""
    from nhvr_b200 import capi
    _case(cuda_dev, 192, 2, 32, 32, capi.ACT_RELU, False, False, (1, 1, 1, 1), capi.HALO_REFLECT)
    _case(cuda_dev, 192, 1, 32, 32, capi.ACT_NONE, True, True, (0, 1, 0, 1), capi.HALO_ZERO)
    _case(cuda_dev, 256, 1, 32, 32, capi.ACT_NONE, True, False, (3, 3, 3, 3), capi.HALO_REFLECT)


def test_in_fused_generator_matches_two_kernel_path(cuda_dev, monkeypatch):
    """The whole temporal generator (10 ResnetBlocks at 128^2 = 130 tiles per image) with and without the fused epilogue."""
    from nhvr_b200.networks import define_G
    torch.manual_seed(3)
    monkeypatch.setenv("NHVR_IN_FUSED", "1")          # plain fp16 engines take the two-kernel path by default
    net = define_G(12, 4, 48, "temporal", n_downsample_global=2, n_blocks_global=10).to(cuda_dev).eval()
    x = torch.rand(2, 12, 512, 512, device=cuda_dev) * 2 - 1
    with torch.no_grad():
        y_fused = net(x).clone()
        eng = next(iter(net._engines.values()))
        assert sum(eng.fused) >= 20, eng.fused
        monkeypatch.setenv("NHVR_IN_FUSED", "0")
        net._engines.clear()
        y_two = net(x).clone()
        eng2 = next(iter(net._engines.values()))
        assert not any(eng2.fused)
    assert (y_fused - y_two).abs().max().item() < 2e-2
