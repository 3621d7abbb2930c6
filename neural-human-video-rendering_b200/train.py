"""Training steps of the reference on the sm_100a kernels, and the data-parallel gradient exchange.

UV-generator pre-train (REF pretrainTrans.sh -> pre_train.py; README.md:68-74): pose -> TransG -> 73-ch output,
objective lambda_UV * masked-L1(UV) + lambda_Prob * CE(part) against DensePose (REF pretrain_start.sh:32-34),
Adam(lr 2e-4, beta1 0.5) as in pix2pixHD.  One process per GPU; gradients are averaged with ONE NCCL
all-reduce over a flat fp32 bucket after backward (the path's only collective).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import losses


class FlatGradBucket:
    """Flat fp32 view over the gradients of a parameter list: one all-reduce per step, no per-tensor launches."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def gather(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def scatter(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)

    def all_reduce_mean(self, group=None) -> None:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.gather()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))
        self.scatter()


class UVPretrainer:
    """configs[1]: UV generator pre-train step (forward, objective, backward, optional DDP all-reduce, Adam)."""

    def __init__(self, netTransG, lr: float = 2e-4, beta1: float = 0.5, lambda_uv: float = 1000.0, lambda_prob: float = 10.0,
                 distributed: bool = False):
        self.net = netTransG
        self.lambda_uv, self.lambda_prob = lambda_uv, lambda_prob
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, betas=(beta1, 0.999))
        self.bucket: Optional[FlatGradBucket] = FlatGradBucket(self.net.parameters()) if distributed else None

    def step(self, pose: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        uvp = self.net(pose)
        loss = losses.uv_prob_objective(uvp, dp_i, dp_uv, self.lambda_uv, self.lambda_prob)
        loss.backward()
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
        self.opt.step()
        return loss.detach()


def synthetic_densepose(N: int, H: int, W: int, device, seed: int = 0):
    """Synthetic pose / DensePose targets of the pre-train shapes (SURVEY §8d cfg 2): piecewise-constant part
    blobs, UV uniform in [0,1], pose = smooth random field in [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randint(0, 25, (N, 1, 8, 8), generator=g).float()
    dp_i = torch.nn.functional.interpolate(low, size=(H, W), mode="nearest")[:, 0].long()
    dp_uv = torch.rand(N, 2, H, W, generator=g)
    pose = torch.tanh(torch.nn.functional.interpolate(torch.randn(N, 3, 16, 16, generator=g), size=(H, W), mode="bilinear") * 2)
    return pose.to(device), dp_i.to(device), dp_uv.to(device)


class RenderTrainer:
    """configs[2]: the end-to-end training step of train.py [REF train_start/pretrain_start.sh:9-37].

    Two consecutive frames per sample (the temporal term needs t-1): frame t-1 is rendered without gradient
    from a zero previous frame, frame t is conditioned on it.  Objectives (pix2pixHD + the reference's lambdas):
      D: 0.5 * (LSGAN(D(cond, fake.detach()), 0) + LSGAN(D(cond, real), 1))
      G: LSGAN(D(cond, fake), 1) + lambda_feat * FM + lambda_L2 * MSE + lambda_UV * UV + lambda_Prob * CE
         + lambda_Temp * L1(out_t - warp(out_{t-1}, flow_inv))
    Adam(2e-4, beta1 0.5) for the G side (three networks, atlas, background image) and for D.  Under
    torch.distributed each side does ONE flat-bucket NCCL all-reduce after its backward.
    """

    def __init__(self, pipe, netD, lr=2e-4, beta1=0.5, lambda_feat=10.0, lambda_l2=500.0, lambda_uv=1000.0, lambda_prob=10.0,
                 lambda_temp=500.0, n_layers_D=3, num_D=2, distributed=False):
        self.pipe, self.netD = pipe, netD
        self.lam = dict(feat=lambda_feat, l2=lambda_l2, uv=lambda_uv, prob=lambda_prob, temp=lambda_temp)
        self.n_layers_D, self.num_D = n_layers_D, num_D
        self.opt_G = torch.optim.Adam(self.pipe.parameters(), lr=lr, betas=(beta1, 0.999))
        self.opt_D = torch.optim.Adam(self.netD.parameters(), lr=lr, betas=(beta1, 0.999))
        self.bucket_G = FlatGradBucket(self.pipe.parameters()) if distributed else None
        self.bucket_D = FlatGradBucket(self.netD.parameters()) if distributed else None

    def step(self, batch: dict) -> dict:
        """pix2pixHD's optimize_parameters order: every discriminator pass of the step (fake detached, real, fake for the
        generator) sees the SAME discriminator weights; the generator side is updated first, then the discriminator."""
        pipe, D, lam = self.pipe, self.netD, self.lam
        pose0, pose1, real1 = batch["pose_prev"], batch["pose"], batch["image"]
        with torch.no_grad():
            r0 = pipe.forward_train(pose0, torch.zeros_like(real1))
        r1 = pipe.forward_train(pose1, r0["out"])
        fake = r1["out"]
        # ---- discriminator passes (no weight update in between)
        pred_fake_d = D(pose1, fake.detach())
        pred_real = D(pose1, real1)
        loss_D = losses.lsgan_diff(pred_fake_d, False, 0.5) + losses.lsgan_diff(pred_real, True, 0.5)
        for p in D.parameters():
            p.requires_grad_(False)
        pred_fake = D(pose1, fake)
        for p in D.parameters():
            p.requires_grad_(True)
        loss_G = (losses.lsgan_diff(pred_fake, True)
                  + losses.feature_matching_diff(pred_fake, [[t.detach() for t in s] for s in pred_real], self.n_layers_D, self.num_D, lam["feat"])
                  + losses.mse_diff(fake, real1, lam["l2"])
                  + losses.uv_prob_objective(r1["uvp"], batch["dp_i"], batch["dp_uv"], lam["uv"], lam["prob"])
                  + losses.temporal_diff(fake, r0["out"], batch["flow_inv"], lam["temp"]))
        # ---- generator side
        self.opt_G.zero_grad(set_to_none=True)
        loss_G.backward()
        if self.bucket_G is not None:
            self.bucket_G.all_reduce_mean()
        self.opt_G.step()
        # ---- discriminator
        self.opt_D.zero_grad(set_to_none=True)
        loss_D.backward()
        if self.bucket_D is not None:
            self.bucket_D.all_reduce_mean()
        self.opt_D.step()
        return {"loss_D": loss_D.detach(), "loss_G": loss_G.detach()}


def synthetic_train_batch(N: int, size: int, device, seed: int = 0) -> dict:
    """Synthetic sample of the end-to-end shapes (SURVEY §8d cfg 3): poses, real image U(-1,1), DensePose IUV, flow N(0, 2 px)."""
    pose, dp_i, dp_uv = synthetic_densepose(N, size, size, device, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    pose_prev = torch.roll(pose, shifts=3, dims=-1)
    image = torch.tanh(torch.nn.functional.interpolate(torch.randn(N, 3, 32, 32, generator=g), size=(size, size), mode="bilinear")).to(device)
    flow = (torch.randn(N, 2, size, size, generator=g) * 2.0).to(device)
    return {"pose_prev": pose_prev, "pose": pose, "image": image, "dp_i": dp_i, "dp_uv": dp_uv, "flow_inv": flow}
