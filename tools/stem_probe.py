"""Stem accuracy on a stick-figure pose map (split precision): raw conv output, statistics and normalised output against fp64."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from nhvr_b200 import capi, ops, pose as posemod

dev = torch.device("cuda", 0)
capi.set_operand_dtype("f16")
kps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keypoints_body25.npy"))[:1]
torch.backends.cudnn.allow_tf32 = False
for shift in (0.0,):
    x = torch.from_numpy(posemod.pose_maps(kps, 512, 6)).to(dev) + shift
    torch.manual_seed(11)
    w = (torch.randn(64, 6, 7, 7) * 0.02).to(dev)
    plan = ops.ConvPlan(capi.CONV, 6, 64, 7, 1, 3, 1, 512, 512, capi.HALO_REFLECT, capi.EPI_RAW_STATS, split3=True)
    xin = ops.P8Buffer(plan.in_desc.copy(), dev)
    ops.pack_nchw([x], xin)
    plan.pack_weights(w)
    raw = ops.P8Buffer(plan.raw_desc(), dev)
    stats = torch.zeros(plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
    if len(sys.argv) > 1:
        ops.stem_stat_shift(w.sum((2, 3)).contiguous(), [x], stats)
    plan.forward(xin, raw.ptr, stats=stats)
    r = ops.unpack_nchw(raw, 64).double()
    ref = F.conv2d(F.pad(x.double(), (3,) * 4, mode="reflect"), w.double())
    d = (r - ref)
    sig = ref.std((2, 3), keepdim=True)
    print("shift %.0f: raw err max %.3e rms %.3e | in sigma units max %.3e rms %.3e | mean^2/var max %.1f"
          % (shift, d.abs().max().item(), d.pow(2).mean().sqrt().item(), (d / sig).abs().max().item(), (d / sig).pow(2).mean().sqrt().item(),
             (ref.mean((2, 3)) ** 2 / ref.var((2, 3), unbiased=False)).max().item()))
    st = stats.view(64, 4)
    n = 512.0 * 512.0
    mean, var = st[:, 0] / n + st[:, 2], st[:, 1] / n - (st[:, 0] / n) ** 2
    mean_r, var_r = ref.mean((2, 3))[0], ref.var((2, 3), unbiased=False)[0]
    print("   stats: mean err / sigma max %.3e | var rel err max %.3e | shift[:4] %s"
          % (((mean - mean_r).abs() / var_r.sqrt()).max().item(), ((var - var_r).abs() / var_r).max().item(), st[:4, 2].tolist()))
    dst = ops.P8Buffer(ops.make_desc(1, plan.Cout8, 512, 512, (1, 1, 1, 1), 0, capi.HALO_REFLECT, hilo=1), dev)
    ops.in_apply(raw, stats, capi.ACT_RELU, dst)
    y = ops.unpack_nchw(dst, 64).double()
    yr = torch.relu(F.instance_norm(ref, eps=1e-5))
    print("   normalised+ReLU: err max %.3e rms %.3e (|y| max %.1f)" % ((y - yr).abs().max().item(), (y - yr).pow(2).mean().sqrt().item(), yr.abs().max().item()))
    y32 = torch.relu(F.instance_norm(F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w), eps=1e-5)).double()
    print("   torch fp32 same op: err max %.3e rms %.3e" % ((y32 - yr).abs().max().item(), (y32 - yr).pow(2).mean().sqrt().item()))
