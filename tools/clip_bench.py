"""Frames/s of RenderPipeline.render_keypoints for B clips of T frames (device keypoints in, device frames out; B <= 2 takes the
two-stage frame pipeline).  usage: python tools/clip_bench.py [B] [T] [strict|balanced|fast] [size]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200.pipeline import RenderPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
prec = sys.argv[3] if len(sys.argv) > 3 else "strict"
SZ = int(sys.argv[4]) if len(sys.argv) > 4 else 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pipe = RenderPipeline(size=SZ, pose_nc=6, precision=prec).to(dev).eval()
kps = torch.rand(B, T, 25, 3, device=dev) * torch.tensor([1024.0, 1024.0, 1.0], device=dev)
out = torch.empty(B, T, 3, SZ, SZ, device=dev)
pipe.render_keypoints(kps[:, :8], SZ, out=out[:, :8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record()
    pipe.render_keypoints(kps, SZ, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("B=%d T=%d %s %d^2 fused=%s: %.1f frames/s (%.3f ms per frame step)" % (B, T, prec, SZ, "off" if os.environ.get("NHVR_NO_IN_FUSED") else "on",
                                                                             B * T / best * 1e3, best / T))
