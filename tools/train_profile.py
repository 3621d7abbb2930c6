"""Per-kernel-class time of the training steps (event-bracketed launches through ops.PROFILE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nhvr_b200 import ops
from nhvr_b200.networks import define_G, define_D
from nhvr_b200.pipeline import RenderPipeline
from nhvr_b200.train import UVPretrainer, RenderTrainer, synthetic_densepose, synthetic_train_batch
from bench import PIPE_KW

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "uv"
if mode == "uv":
    net = define_G(3, 73, 64, "translate", 2, 5)
    tr = UVPretrainer(net)
    data = synthetic_densepose(16, 256, 256, dev)
    step = lambda: tr.step(*data)
else:
    pipe = RenderPipeline(**PIPE_KW).to(dev)
    netD = define_D(PIPE_KW["pose_nc"] + 3, 64, 3, "instance", False, 2, True)
    tr = RenderTrainer(pipe, netD)
    batch = synthetic_train_batch(8, 512, dev)
    if "--stick" in sys.argv:      # the reference's real input: BODY_25 stick figures rasterised from the bundled keypoints (~98 % flat)
        import numpy as np
        from nhvr_b200 import pose as posemod
        kps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keypoints_body25.npy"))
        maps = torch.from_numpy(posemod.pose_maps(kps[:16], 512, 3)).to(dev)
        batch["pose_prev"], batch["pose"] = maps[0::2].contiguous(), maps[1::2].contiguous()
    z = torch.zeros(8, PIPE_KW["pose_nc"] - 3, 512, 512, device=dev)            # the zero LaplaceProj channels of --use_laplace
    batch["pose"], batch["pose_prev"] = torch.cat([batch["pose"], z], 1), torch.cat([batch["pose_prev"], z], 1)
    step = lambda: tr.step(batch)
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
print("step ms (unprofiled):", e0.elapsed_time(e1) / 3)
ops.PROFILE = []
step()
torch.cuda.synchronize()
recs, ops.PROFILE = ops.PROFILE, None
agg = {}
for kind, work, a, b in recs:
    d = agg.setdefault(kind, [0.0, 0.0, 0]); d[0] += work; d[1] += a.elapsed_time(b); d[2] += 1
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rate = v[0] / (v[1] * 1e-3)
    print("%-12s n=%4d  %8.2f ms  %5.1f%%  %s" % (k, v[2], v[1], 100 * v[1] / tot, ("%.0f TFLOP/s" % (rate / 1e12)) if k in ("conv", "wgrad") else ("%.0f GB/s" % (rate / 1e9))))
print("sum of bracketed launches: %.2f ms" % tot)
if "--launches" in sys.argv:
    for i, (kind, work, a, b) in enumerate(recs):
        ms = a.elapsed_time(b)
        if kind in ("conv", "wgrad"):
            print("%3d %-8s %8.1f us  work=%.3e  %7.1f TFLOP/s" % (i, kind, ms * 1e3, work, work / ms / 1e9))
        else:
            print("%3d %-8s %8.1f us  work=%.3e  %7.1f GB/s" % (i, kind, ms * 1e3, work, work / ms / 1e6))
