"""One end-to-end training step (configs[2]: 512^2, batch 8) inside a cudaProfiler range, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` (tools/ncu_summarise.py launches <csv> gives the per-kernel share)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bench import PIPE_KW
from nhvr_b200.networks import define_D
from nhvr_b200.pipeline import RenderPipeline
from nhvr_b200.train import RenderTrainer, synthetic_train_batch

dev = torch.device("cuda", 0)
torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "uv":            # configs[1]: UV generator pre-train, 256^2, batch 16
    from nhvr_b200.networks import define_G
    from nhvr_b200.train import UVPretrainer, synthetic_densepose
    tr_uv = UVPretrainer(define_G(3, 73, 64, "translate", 2, 5))
    data = synthetic_densepose(16, 256, 256, dev)
    step = lambda: tr_uv.step(*data)
else:
    pipe = RenderPipeline(**PIPE_KW).to(dev)
    netD = define_D(PIPE_KW["pose_nc"] + 3, 64, 3, "instance", False, 2, True)
    tr = RenderTrainer(pipe, netD)
    bt = synthetic_train_batch(8, 512, dev)
    z = torch.zeros(8, PIPE_KW["pose_nc"] - 3, 512, 512, device=dev)
    bt["pose"], bt["pose_prev"] = torch.cat([bt["pose"], z], 1), torch.cat([bt["pose_prev"], z], 1)
    step = lambda: tr.step(bt)
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
