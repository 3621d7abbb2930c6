"""Texture-generator pre-train entry point named by pre_train_tex.sh:1 - parses the reference's flags verbatim
(--input_nc 81 --loadSize 200 --n_blocks_global 5 --ngf_global 64 --TexG part --part_texture_path --pose_texture_path
--lapalce_path ...) and runs the pre-train step on the sm_100a kernels.

What the reference's pre_train_tex.py does is not in the mount; SPEC: the texture generator is a
GlobalGenerator(--input_nc -> 24 parts x tex_nc channels, --ngf_global, --n_downsample_global, --n_blocks_global) working at
the part-texture resolution --loadSize (200 = the atlas part size), trained with an L1 objective against the part
textures (--part_texture_path, e.g. unfold_texture.py's output); 81 input channels = 27 x 3: the 24 part textures' Laplace
texture planes plus the pose image (--pose_texture_path / --lapalce_path).  The dataset readers are out of scope
(SURVEY 2): `--synthetic_steps N` exercises the step on synthetic tensors of those shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch

from nhvr_b200 import capi, losses
from nhvr_b200.checkpoint import net_path
from nhvr_b200.networks import define_G
from nhvr_b200.options import TrainOptions
from nhvr_b200.train import FlatGradBucket


def main(argv=None):
    to = TrainOptions()
    to.initialize()
    to.parser.add_argument("--synthetic_steps", type=int, default=0, help="train on synthetic tensors for N steps")
    opt = to.parse(argv)
    capi.require_device()
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0] if opt.gpu_ids[0] < torch.cuda.device_count() else 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if opt.synthetic_steps <= 0:
        for pth in (opt.part_texture_path, opt.pose_texture_path):
            if not pth or not os.path.isdir(pth):
                raise SystemExit("pre_train_tex.py: dataset directory %r not found; the reference's dataset readers are out of "
                                 "scope - use --synthetic_steps N to exercise the training step" % pth)
        raise SystemExit("pre_train_tex.py: real-data loading is not built (SURVEY 2: data/ is out of scope)")
    out_nc = 24 * opt.tex_nc
    torch.manual_seed(0)
    net = define_G(opt.input_nc, out_nc, opt.ngf_global, "global", opt.n_downsample_global, opt.n_blocks_global, gpu_ids=[local])
    optim = torch.optim.Adam(net.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999))
    bucket = FlatGradBucket(net.parameters()) if world > 1 else None
    g = torch.Generator().manual_seed(int(os.environ.get("RANK", 0)))
    S = opt.loadSize
    x = torch.tanh(torch.nn.functional.interpolate(torch.randn(opt.batchSize, opt.input_nc, 10, 10, generator=g), size=S, mode="bilinear")).to(dev)
    target = torch.tanh(torch.nn.functional.interpolate(torch.randn(opt.batchSize, out_nc, 10, 10, generator=g), size=S, mode="bilinear")).to(dev)
    for it in range(opt.synthetic_steps):
        optim.zero_grad(set_to_none=True)
        loss = losses.l1_diff(net(x), target, 1.0)
        loss.backward()
        if bucket is not None:
            bucket.all_reduce_mean()
        optim.step()
        if it % max(1, opt.print_freq // 10) == 0 or it == opt.synthetic_steps - 1:
            print("[pre_train_tex.py] step %d L1 %.4f" % (it, loss.item()))
    if int(os.environ.get("RANK", 0)) == 0:
        save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        os.makedirs(save_dir, exist_ok=True)
        torch.save(net.state_dict(), net_path(save_dir, "latest", "TexG"))
        print("[pre_train_tex.py] saved", net_path(save_dir, "latest", "TexG"))


if __name__ == "__main__":
    main()
