"""End-to-end training entry point named by train_start/pretrain_start.sh:9.  Parses the reference's flags verbatim
and runs the training step on the sm_100a kernels (RenderTrainer: D step + G step, Adam, one NCCL all-reduce per
side under torchrun).  The reference's dataset readers (image / mask / DensePose / flow directories, README.md:39-62)
are out of scope: `--synthetic_steps N` exercises the step on synthetic samples of the same shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch

from nhvr_b200 import capi
from nhvr_b200.checkpoint import load_pipeline, save_pipeline, net_path
from nhvr_b200.networks import define_D
from nhvr_b200.options import TrainOptions, pipeline_kwargs
from nhvr_b200.pipeline import RenderPipeline
from nhvr_b200.train import RenderTrainer, synthetic_train_batch


def main(argv=None):
    to = TrainOptions()
    to.initialize()
    to.parser.add_argument("--synthetic_steps", type=int, default=0, help="train on synthetic samples for N steps")
    opt = to.parse(argv)
    capi.require_device()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if opt.synthetic_steps <= 0:
        for pth in (opt.pose_path, opt.img_path, opt.densepose_path, opt.mask_path, opt.flow_inv_path):
            if not pth or not os.path.isdir(pth):
                raise SystemExit("train.py: dataset directory %r not found; the reference's dataset readers are out of scope — "
                                 "use --synthetic_steps N to exercise the training step" % pth)
        raise SystemExit("train.py: real-data loading is not built (SURVEY §2: data/ is out of scope)")
    torch.manual_seed(0)                                     # identical initial weights on every rank
    pipe = RenderPipeline(**pipeline_kwargs(opt)).to(dev)
    if opt.load_pretrain_TransG:
        p = net_path(opt.load_pretrain_TransG, opt.which_epoch_TransG, "TransG")
        if os.path.isfile(p):
            sd = torch.load(p, map_location="cpu")
            # pretrainTrans.sh trains the UV generator on the 3-channel pose map (--input_nc 3, no --use_laplace) while
            # pretrain_start.sh feeds pose + LaplaceProj (--use_laplace): widen the stem with zero weights for the extra
            # input channels - the same function on the first channels
            w = sd.get("model.1.weight")
            if w is not None and w.shape[1] < opt.pose_nc:
                sd["model.1.weight"] = torch.cat([w, torch.zeros(w.shape[0], opt.pose_nc - w.shape[1], *w.shape[2:])], 1)
            pipe.netTransG.load_state_dict(sd)
            print("[train.py] loaded", p)
        else:
            print("[train.py] --load_pretrain_TransG: %s not found, UV generator starts from random init" % p)
    if opt.continue_train:
        load_pipeline(pipe, os.path.join(opt.checkpoints_dir, opt.name), opt.which_epoch)
    netD = define_D(opt.pose_nc + 3, opt.ndf, opt.n_layers_D, "instance", False, opt.num_D, not opt.no_ganFeat_loss, gpu_ids=[local])
    vgg = None
    if not opt.no_vgg_loss:                      # pix2pixHD default: VGGLoss on; it needs the ImageNet vgg19 weights
        if opt.vgg_weights and os.path.isfile(opt.vgg_weights):
            from nhvr_b200.networks import Vgg19B200
            vgg = Vgg19B200()
            sd = torch.load(opt.vgg_weights, map_location="cpu")
            vgg.load_state_dict({k: v for k, v in sd.items() if k.startswith("features.") and int(k.split(".")[1]) <= 28})
            vgg = vgg.to(dev)
            print("[train.py] perceptual loss: loaded", opt.vgg_weights)
        else:
            print("[train.py] perceptual (VGG) loss skipped: pass --vgg_weights <vgg19 state_dict> (no ImageNet weights offline)")
    trainer = RenderTrainer(pipe, netD, lr=opt.lr, beta1=opt.beta1, lambda_feat=opt.lambda_feat, lambda_l2=opt.lambda_L2,
                            lambda_uv=opt.lambda_UV, lambda_prob=opt.lambda_Prob, lambda_temp=opt.lambda_Temp,
                            n_layers_D=opt.n_layers_D, num_D=opt.num_D, distributed=world > 1, vgg=vgg)
    batch = synthetic_train_batch(opt.batchSize, opt.loadSize, dev, seed=rank)
    if opt.pose_nc > 3:
        z = torch.zeros(opt.batchSize, opt.pose_nc - 3, opt.loadSize, opt.loadSize, device=dev)
        batch["pose"] = torch.cat([batch["pose"], z], 1)
        batch["pose_prev"] = torch.cat([batch["pose_prev"], z], 1)
    for it in range(opt.synthetic_steps):
        out = trainer.step(batch)
        if it % max(1, opt.print_freq // 10) == 0 or it == opt.synthetic_steps - 1:
            print("[train.py] step %d loss_G %.4f loss_D %.4f" % (it, out["loss_G"].item(), out["loss_D"].item()))
    if rank == 0:
        save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        save_pipeline(pipe, save_dir, "latest")
        torch.save(netD.state_dict(), net_path(save_dir, "latest", "D"))
        print("[train.py] saved checkpoints under", save_dir)


if __name__ == "__main__":
    main()
