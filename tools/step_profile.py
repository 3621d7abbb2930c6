"""Per-launch times of ONE rendering step (event-bracketed, un-graphed): kind, work, ms, rate.
usage: python tools/step_profile.py [clips]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200 import ops
from nhvr_b200.pipeline import RenderPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pipe = RenderPipeline().to(dev).eval()
eager = pipe.step_graph(B, 512, 512, use_graph=False)
poses = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
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
print(f"{n} launches per step, clips={B}")
for i in range(n):
    kind, work = recs[i][0], recs[i][1]
    ms = sum(recs[i + r * n][2].elapsed_time(recs[i + r * n][3]) for r in range(REP)) / REP
    unit = "TFLOP/s" if kind == "conv" else "GB/s"
    rate = work / ms / (1e9 if kind == "conv" else 1e6)
    tot[kind] = tot.get(kind, 0.0) + ms
    print(f"{i:3d} {kind:10s} work={work:.3e} {ms*1e3:8.1f} us  {rate:8.1f} {unit}")
print({k: round(v, 3) for k, v in tot.items()}, "sum ms", round(sum(tot.values()), 3))
