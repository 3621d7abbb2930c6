"""Frames/s of RenderPipeline.render_keypoints against the number of lock-step clips B (one process, one pipeline per precision).
A 128 x 128 ResnetBlock layer has 130 tiles per clip and the GPU 296 resident CTA slots, so the tile count of a launch is
130 * B: B = 8 is 3.51 rounds of CTAs, B = 9 is 3.95 (wave quantisation of the last round).
usage: python tools/clip_sweep.py [strict|balanced|fast] [T] [B ...]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200.pipeline import RenderPipeline

prec = sys.argv[1] if len(sys.argv) > 1 else "strict"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
Bs = [int(a) for a in sys.argv[3:]] or [6, 7, 8, 9, 10, 11, 12, 16, 18]
SZ = 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pipe = RenderPipeline(size=SZ, pose_nc=6, precision=prec).to(dev).eval()
for B in Bs:
    kps = torch.rand(B, T, 25, 3, device=dev) * torch.tensor([1024.0, 1024.0, 1.0], device=dev)
    out = torch.empty(B, T, 3, SZ, SZ, device=dev)
    pipe.render_keypoints(kps[:, :4], SZ, out=out[:, :4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        pipe.render_keypoints(kps, SZ, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s B=%2d T=%d: %7.1f frames/s  %.3f ms per frame step  %.3f ms per frame" % (prec, B, T, B * T / best * 1e3, best / T, best / T / B), flush=True)
    pipe._graphs.clear()
    for net in (pipe.netTransG, pipe.netG):
        getattr(net, "_engines", {}).clear()
    del kps, out
    torch.cuda.empty_cache()
