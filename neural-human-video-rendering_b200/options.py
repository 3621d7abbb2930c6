"""Option system: every flag the reference's launch scripts pass, verbatim.

The reference's ``options/`` package is absent from the mount; flag NAMES and the values seen are
pinned by the five scripts [REF test_start/start.sh:6-28; train_start/pretrain_start.sh:9-45;
pretrainTrans.sh:1-16; pre_train_tex.sh:1-23].  Upstream pix2pixHD flags that the scripts rely on
implicitly (defaults) follow public pix2pixHD ``options/base_options.py`` / ``train_options.py`` /
``test_options.py`` [SURVEY Appendix C].  Unknown flags are an error, as in argparse upstream.
"""
from __future__ import annotations

import argparse
import os


class BaseOptions:
    def __init__(self):
        self.parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        self.initialized = False
        self.isTrain = False

    def initialize(self):
        p = self.parser
        # experiment specifics (pix2pixHD)
        p.add_argument("--name", type=str, default="label2city", help="experiment dir under checkpoints_dir")
        p.add_argument("--gpu_ids", type=str, default="0", help="gpu ids: e.g. 0  0,1,2. -1 is rejected: no CPU path")
        p.add_argument("--checkpoints_dir", type=str, default="./checkpoints")
        p.add_argument("--model", type=str, default="pix2pixHD")
        p.add_argument("--norm", type=str, default="instance")
        p.add_argument("--use_dropout", action="store_true")
        p.add_argument("--data_type", default=32, type=int, choices=[8, 16, 32])
        p.add_argument("--verbose", action="store_true", default=False)
        p.add_argument("--fp16", action="store_true", default=False)
        p.add_argument("--local_rank", type=int, default=0)
        # input/output sizes
        p.add_argument("--batchSize", type=int, default=1)
        p.add_argument("--loadSize", type=int, default=512, help="REF start.sh:25 (512); pre_train_tex.sh:19 (200)")
        p.add_argument("--fineSize", type=int, default=512)
        p.add_argument("--label_nc", type=int, default=0)
        p.add_argument("--input_nc", type=int, default=3, help="pose-map channels [REF start.sh:24]; 81 for tex pre-train")
        p.add_argument("--output_nc", type=int, default=3)
        # data
        p.add_argument("--dataroot", type=str, default="./datasets/")
        p.add_argument("--resize_or_crop", type=str, default="resize")
        p.add_argument("--serial_batches", action="store_true")
        p.add_argument("--no_flip", action="store_true")
        p.add_argument("--nThreads", default=2, type=int)
        p.add_argument("--max_dataset_size", type=int, default=float("inf"))
        p.add_argument("--data_ratio", type=float, default=0.9, help="train/val split [REF pretrain_start.sh:36]")
        # paths (custom to the reference)
        p.add_argument("--pose_path", type=str, default="./keypoints", help="[REF start.sh:9]")
        p.add_argument("--pose_tgt_path", type=str, default="", help="target person's openpose_json [REF start.sh:10]")
        p.add_argument("--mask_path", type=str, default="")
        p.add_argument("--img_path", type=str, default="")
        p.add_argument("--densepose_path", type=str, default="")
        p.add_argument("--bg_path", type=str, default="", help="bg.jpg [REF start.sh:12]")
        p.add_argument("--texture_path", type=str, default="", help="texture.jpg [REF start.sh:13]")
        p.add_argument("--flow_path", type=str, default="")
        p.add_argument("--flow_inv_path", type=str, default="")
        p.add_argument("--lapalce_path", type=str, default="", help="(sic) [REF pre_train_tex.sh:6]")
        p.add_argument("--laplace_path", type=str, default="")
        p.add_argument("--part_texture_path", type=str, default="")
        p.add_argument("--pose_texture_path", type=str, default="")
        # display
        p.add_argument("--display_winsize", type=int, default=512)
        p.add_argument("--tf_log", action="store_true")
        # generator
        p.add_argument("--netG", type=str, default="global")
        p.add_argument("--ngf", type=int, default=64)
        p.add_argument("--ngf_global", type=int, default=48, help="[REF start.sh:17]")
        p.add_argument("--n_downsample_global", type=int, default=2, help="[REF start.sh:15]")
        p.add_argument("--n_blocks_global", type=int, default=10, help="[REF start.sh:16]")
        p.add_argument("--n_blocks_local", type=int, default=3)
        p.add_argument("--n_local_enhancers", type=int, default=1)
        p.add_argument("--niter_fix_global", type=int, default=0)
        p.add_argument("--n_downsample_bg", type=int, default=2, help="[REF start.sh:20]")
        p.add_argument("--n_blocks_bg", type=int, default=2, help="[REF start.sh:21]")
        p.add_argument("--ngf_bg", type=int, default=48, help="SPEC D10")
        p.add_argument("--n_blocks_translate", type=int, default=5, help="[REF pretrainTrans.sh:13]")
        p.add_argument("--ngf_translate", type=int, default=64, help="SPEC D4")
        p.add_argument("--n_downsample_translate", type=int, default=2, help="SPEC D4")
        p.add_argument("--TexG", type=str, default="part", help="[REF start.sh:14]")
        p.add_argument("--tex_nc", type=int, default=3, help="texture channels, SPEC D2")
        p.add_argument("--atlas_size", type=int, default=200, help="part texture size [REF pre_train_tex.sh:19]")
        p.add_argument("--use_mask_texture", action="store_true", help="[REF start.sh:18]")
        p.add_argument("--use_laplace", action="store_true", help="[REF start.sh:11]")
        p.add_argument("--pose_plus_laplace", action="store_true", help="[REF start.sh:19]")
        # instance-wise features (pix2pixHD)
        p.add_argument("--no_instance", action="store_true")
        p.add_argument("--instance_feat", action="store_true", help="[REF start.sh:23]; accepted, unused (SURVEY App. A)")
        p.add_argument("--label_feat", action="store_true")
        p.add_argument("--feat_num", type=int, default=3)
        p.add_argument("--load_features", action="store_true")
        p.add_argument("--n_downsample_E", type=int, default=4)
        p.add_argument("--nef", type=int, default=16)
        p.add_argument("--n_clusters", type=int, default=10)
        # B200 additions (not reference flags)
        p.add_argument("--clips_in_flight", type=int, default=1, help="independent clips advanced in lock-step per GPU")
        p.add_argument("--precision", type=str, default="strict", choices=["strict", "strict2", "balanced", "fast"],
                       help="inference operand precision preset (pipeline.RenderPipeline)")
        self.initialized = True

    def parse(self, args=None, save=False):
        if not self.initialized:
            self.initialize()
        import sys as _sys
        argv = list(_sys.argv[1:] if args is None else args)
        # pretrainTrans.sh:16 ends with a line-continuation backslash at EOF, which bash passes on as a
        # literal "\\" argument; drop it so the script runs unmodified
        argv = [a for a in argv if a.strip() != "\\"]
        self.opt = self.parser.parse_args(argv)
        self.opt.isTrain = self.isTrain
        ids = [int(s) for s in str(self.opt.gpu_ids).split(",") if s.strip() != ""]
        self.opt.gpu_ids = [i for i in ids if i >= 0]
        if not self.opt.gpu_ids:
            raise SystemExit("--gpu_ids -1: this build has no CPU path (sm_100a kernels only)")
        # pose channels actually fed to the networks (SPEC D1 / D12): --use_laplace alone adds the 3 LaplaceProj channels.
        # train_start/pretrain_start.sh passes --use_laplace without --pose_plus_laplace, test_start/start.sh passes both:
        # deriving the count from --use_laplace keeps a checkpoint written by train.py loadable by test.py
        self.opt.pose_nc = self.opt.input_nc + (3 if self.opt.use_laplace else 0)
        if save:
            expr_dir = os.path.join(self.opt.checkpoints_dir, self.opt.name)
            os.makedirs(expr_dir, exist_ok=True)
            with open(os.path.join(expr_dir, "opt.txt"), "wt") as f:
                f.write("------------ Options -------------\n")
                for k, v in sorted(vars(self.opt).items()):
                    f.write("%s: %s\n" % (str(k), str(v)))
                f.write("-------------- End ----------------\n")
        return self.opt


class TestOptions(BaseOptions):
    def initialize(self):
        BaseOptions.initialize(self)
        p = self.parser
        p.add_argument("--ntest", type=int, default=float("inf"))
        p.add_argument("--results_dir", type=str, default="./results/", help="[REF start.sh:27]")
        p.add_argument("--aspect_ratio", type=float, default=1.0)
        p.add_argument("--phase", type=str, default="test")
        p.add_argument("--which_epoch", type=str, default="latest", help="[REF start.sh:28]")
        p.add_argument("--how_many", type=int, default=100000)
        p.add_argument("--cluster_path", type=str, default="features_clustered_010.npy")
        p.add_argument("--use_encoded_image", action="store_true")
        p.add_argument("--export_onnx", type=str)
        p.add_argument("--engine", type=str)
        p.add_argument("--onnx", type=str)
        self.isTrain = False


class TrainOptions(BaseOptions):
    def initialize(self):
        BaseOptions.initialize(self)
        p = self.parser
        p.add_argument("--display_freq", type=int, default=100)
        p.add_argument("--print_freq", type=int, default=100)
        p.add_argument("--save_latest_freq", type=int, default=1000)
        p.add_argument("--save_epoch_freq", type=int, default=10, help="[REF pretrain_start.sh:35]")
        p.add_argument("--no_html", action="store_true")
        p.add_argument("--debug", action="store_true")
        p.add_argument("--continue_train", action="store_true")
        p.add_argument("--load_pretrain", type=str, default="")
        p.add_argument("--load_pretrain_TransG", type=str, default="", help="[REF pretrain_start.sh:29]")
        p.add_argument("--which_epoch", type=str, default="latest")
        p.add_argument("--which_epoch_TransG", type=str, default="latest", help="[REF pretrain_start.sh:30]")
        p.add_argument("--phase", type=str, default="train")
        p.add_argument("--niter", type=int, default=100)
        p.add_argument("--niter_decay", type=int, default=100)
        p.add_argument("--beta1", type=float, default=0.5)
        p.add_argument("--lr", type=float, default=0.0002)
        p.add_argument("--num_D", type=int, default=2)
        p.add_argument("--n_layers_D", type=int, default=3)
        p.add_argument("--ndf", type=int, default=64)
        p.add_argument("--lambda_feat", type=float, default=10.0)
        p.add_argument("--no_ganFeat_loss", action="store_true")
        p.add_argument("--no_vgg_loss", action="store_true")
        p.add_argument("--vgg_weights", type=str, default="",
                       help="torchvision vgg19 state_dict (features.<idx>.*); without it the perceptual loss is skipped (no ImageNet weights offline)")
        p.add_argument("--no_lsgan", action="store_true")
        p.add_argument("--pool_size", type=int, default=0)
        p.add_argument("--lambda_L2", type=float, default=500.0, help="[REF pretrain_start.sh:31]")
        p.add_argument("--lambda_UV", type=float, default=1000.0, help="[REF pretrain_start.sh:32]")
        p.add_argument("--lambda_Prob", type=float, default=10.0, help="[REF pretrain_start.sh:33]")
        p.add_argument("--lambda_Temp", type=float, default=500.0, help="[REF pretrain_start.sh:37]")
        p.add_argument("--use_densepose_loss", action="store_true", help="[REF pretrain_start.sh:34]")
        self.isTrain = True


def pipeline_kwargs(opt) -> dict:
    """Map parsed reference flags onto RenderPipeline's constructor."""
    return dict(pose_nc=opt.pose_nc, tex_nc=opt.tex_nc, size=opt.loadSize, atlas_size=opt.atlas_size,
                ngf_global=opt.ngf_global, n_downsample_global=opt.n_downsample_global,
                n_blocks_global=opt.n_blocks_global, ngf_translate=opt.ngf_translate,
                n_downsample_translate=opt.n_downsample_translate, n_blocks_translate=opt.n_blocks_translate,
                ngf_bg=opt.ngf_bg, n_downsample_bg=opt.n_downsample_bg, n_blocks_bg=opt.n_blocks_bg,
                use_mask_texture=opt.use_mask_texture)
