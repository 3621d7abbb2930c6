"""One frame step of the bench workload inside a cudaProfiler range (for `ncu --profile-from-start off`): 8 lock-step clips
at 512^2, un-graphed so that every kernel is a separate launch.  usage: python tools/ncu_step.py [precision] [clips]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from bench import PIPE_KW, SIZE, synthetic_keypoints
from nhvr_b200 import ops
from nhvr_b200.pipeline import RenderPipeline

prec = sys.argv[1] if len(sys.argv) > 1 else "strict"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pipe = RenderPipeline(**PIPE_KW, precision=prec).to(dev)
step = pipe.step_graph(B, SIZE, SIZE, use_graph=False, pipelined=False)
kps = synthetic_keypoints(B, 4).to(dev)
for t in range(2):
    ops.pose_rasterize(kps[:, t].contiguous(), SIZE, pipe.pose_nc, out=step.pose); step.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.pose_rasterize(kps[:, 2].contiguous(), SIZE, pipe.pose_nc, out=step.pose); step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
