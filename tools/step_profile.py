"""Per-launch times of ONE rendering step (event-bracketed, un-graphed): kind, work, ms, rate.
usage: python tools/step_profile.py [clips] [uv_precision f16|split3] [size] [g_precision f16|split3]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200 import ops
from nhvr_b200.pipeline import RenderPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
UVP = sys.argv[2] if len(sys.argv) > 2 else "split3"
SZ = int(sys.argv[3]) if len(sys.argv) > 3 else 512
GP = sys.argv[4] if len(sys.argv) > 4 else "f16"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pipe = RenderPipeline(size=SZ, uv_precision=UVP, g_precision=GP).to(dev).eval()
eager = pipe.step_graph(B, SZ, SZ, use_graph=False)
poses = torch.rand(B, 3, SZ, SZ, device=dev) * 2 - 1
for _ in range(3):
    eager.pose.copy_(poses); eager.run()
torch.cuda.synchronize()
tot = {}
REP = 5
ops.PROFILE = []
for _ in range(REP):
    eager.pose.copy_(poses); eager.run()
torch.cuda.synchronize()
recs, ops.PROFILE = ops.PROFILE, None
n = len(recs) // REP
print(f"{n} launches per step, clips={B}, uv_precision={UVP}, g_precision={GP}, size={SZ}")
for i in range(n):
    kind, work = recs[i][0], recs[i][1]
    ms = sum(recs[i + r * n][2].elapsed_time(recs[i + r * n][3]) for r in range(REP)) / REP
    unit = "TFLOP/s" if kind.startswith("conv") else "GB/s"
    rate = work / ms / (1e9 if kind.startswith("conv") else 1e6)
    tot[kind.split(":")[0]] = tot.get(kind.split(":")[0], 0.0) + ms
    print(f"{i:3d} {kind:28s} work={work:.3e} {ms*1e3:8.1f} us  {rate:8.1f} {unit}")
print({k: round(v, 3) for k, v in tot.items()}, "sum ms", round(sum(tot.values()), 3))

# whole step under a CUDA graph
g = pipe.step_graph(B, SZ, SZ, use_graph=True)
for _ in range(3):
    g.pose.copy_(poses); g.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("graph step %.3f ms -> %.1f frames/s (no L2 flush)" % (ms, B * 1000.0 / ms))
