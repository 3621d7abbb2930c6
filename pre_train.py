"""UV-generator pre-train entry point named by pretrainTrans.sh:1.  Parses the reference's flags verbatim and
runs the pre-train step on the sm_100a kernels (forward, lambda_UV/lambda_Prob objective, backward, Adam).
The reference's dataset readers (OpenPose JSON + DensePose + mask directories, README.md:74) are out of scope
(SURVEY §2); without `--synthetic_steps` and with unreadable data paths this driver says so instead of guessing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch

from nhvr_b200 import capi
from nhvr_b200.checkpoint import net_path
from nhvr_b200.networks import define_G
from nhvr_b200.options import TrainOptions
from nhvr_b200.train import UVPretrainer, synthetic_densepose


def main(argv=None):
    to = TrainOptions()
    to.initialize()
    to.parser.add_argument("--synthetic_steps", type=int, default=0, help="train on synthetic DensePose targets for N steps")
    opt = to.parse(argv)
    capi.require_device()
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0]))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if opt.synthetic_steps <= 0:
        for pth in (opt.pose_path, opt.densepose_path, opt.mask_path):
            if not pth or not os.path.isdir(pth):
                raise SystemExit("pre_train.py: dataset directory %r not found; the reference's dataset readers are out of "
                                 "scope — use --synthetic_steps N to exercise the training step" % pth)
        raise SystemExit("pre_train.py: real-data loading is not built (SURVEY §2: data/ is out of scope)")
    net = define_G(opt.pose_nc, 73, opt.ngf_translate, "translate", opt.n_downsample_translate, opt.n_blocks_translate,
                   gpu_ids=[local])
    trainer = UVPretrainer(net, lr=opt.lr, beta1=opt.beta1, lambda_uv=opt.lambda_UV, lambda_prob=opt.lambda_Prob,
                           distributed=world > 1)
    pose, dp_i, dp_uv = synthetic_densepose(opt.batchSize, opt.loadSize, opt.loadSize, torch.device("cuda", local),
                                            seed=int(os.environ.get("RANK", 0)))
    if opt.pose_nc > 3:
        pose = torch.cat([pose, torch.zeros(pose.shape[0], opt.pose_nc - 3, *pose.shape[2:], device=pose.device)], 1)
    for it in range(opt.synthetic_steps):
        loss = trainer.step(pose, dp_i, dp_uv)
        if it % max(1, opt.print_freq // 10) == 0 or it == opt.synthetic_steps - 1:
            print("[pre_train.py] step %d loss %.4f" % (it, loss.item()))
    save_dir = os.path.join(opt.checkpoints_dir, opt.name)
    if int(os.environ.get("RANK", 0)) == 0:
        os.makedirs(save_dir, exist_ok=True)
        torch.save(net.state_dict(), net_path(save_dir, "latest", "TransG"))
        print("[pre_train.py] saved", net_path(save_dir, "latest", "TransG"))


if __name__ == "__main__":
    main()
