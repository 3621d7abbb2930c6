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


class ParamBucket:
    """Parameters, gradients and Adam state of one network as flat fp32 buffers.

    * every parameter's storage is re-pointed into ``flat_p`` and its ``.grad`` into ``flat_g`` (views): the weight-gradient
      kernels accumulate straight into the bucket (networks._ChainEngine.backward), nothing is gathered or scattered;
    * ``all_reduce_async`` launches ONE NCCL all-reduce of ``flat_g`` as soon as the network's backward has finished
      (networks: ``after_backward`` hook) so that it runs under the remaining backward of the other networks;
    * ``adam_step`` is one native launch (nhvr_adam_step) on CUDA; the 1 / world_size of the mean is folded into it.
    On CPU (the gloo tests of the host logic) the same class falls back to torch arithmetic for the update."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8,
                 owners: Iterable[torch.nn.Module] = ()):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = (n + 3) // 4 * 4
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            p._nhvr_direct_grad = True
            off += k
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.steps = 0
        self.works: list = []
        self._tail_from = None
        self.world = 1
        self.owners = [m for o in owners for m in o.modules()]

    def zero_grad(self) -> None:
        self.flat_g.zero_()

    def _offset_of_param(self, index: int) -> int:
        return sum(p.numel() for p in self.params[:index])

    def all_reduce_tail_async(self, first_param: int, group=None) -> None:
        """First of two buckets: the gradients of parameters[first_param:] are complete (the chain's backward has passed the layer
        that owns them - networks._ChainEngine ``mid_backward`` hook) and are reduced under the rest of the backward."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or self._tail_from is not None:
            return
        off = self._offset_of_param(first_param)
        if off <= 0 or off >= self.n:
            return
        self.world = dist.get_world_size(group)
        self._tail_from = off
        self.works.append(dist.all_reduce(self.flat_g[off:], op=dist.ReduceOp.SUM, group=group, async_op=True))

    def all_reduce_async(self, group=None) -> None:
        """The bucket (or, after all_reduce_tail_async, its remaining head) - launched when the network's backward has finished."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.world = dist.get_world_size(group)
        buf = self.flat_g if self._tail_from is None else self.flat_g[:self._tail_from]
        self.works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True))

    def wait(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []
        self._tail_from = None

    def adam_step(self) -> None:
        self.wait()
        self.steps += 1
        gs = 1.0 / self.world
        if self.flat_p.is_cuda:
            from . import capi
            capi.check(capi.load().nhvr_adam_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.n,
                                                  self.lr, self.b1, self.b2, self.eps, gs, self.steps, capi.stream_ptr()), "nhvr_adam_step")
        else:                                                   # host-logic tests only
            g = self.flat_g * gs
            self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1, bc2 = 1 - self.b1 ** self.steps, 1 - self.b2 ** self.steps
            self.flat_p.addcdiv_(self.m, self.v.sqrt() / bc2 ** 0.5 + self.eps, value=-self.lr / bc1)
        self.world = 1
        for mod in self.owners:                                 # the kernel wrote the storage behind the parameters' backs:
            mod.ext_version = getattr(mod, "ext_version", 0) + 1    # packed-weight caches key on this counter too


class UVPretrainer:
    """configs[1]: UV generator pre-train step (forward, objective, backward, optional DDP all-reduce, Adam)."""

    def __init__(self, netTransG, lr: float = 2e-4, beta1: float = 0.5, lambda_uv: float = 1000.0, lambda_prob: float = 10.0,
                 distributed: bool = False):
        self.net = netTransG
        self.lambda_uv, self.lambda_prob = lambda_uv, lambda_prob
        self.bucket = ParamBucket(self.net.parameters(), lr, beta1, owners=[self.net])
        self.distributed = distributed
        if distributed:
            self.net._after_backward = self.bucket.all_reduce_async
            self.net._mid_backward = self.bucket.all_reduce_tail_async      # two buckets: the second half of the layers goes first

    def step(self, pose: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor) -> torch.Tensor:
        self.bucket.zero_grad()
        uvp = self.net(pose)
        loss = losses.uv_prob_objective(uvp, dp_i, dp_uv, self.lambda_uv, self.lambda_prob)
        loss.backward()
        self.bucket.adam_step()
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
                 lambda_temp=500.0, n_layers_D=3, num_D=2, distributed=False, vgg=None):
        """vgg: a networks.Vgg19B200 (with ImageNet weights loaded) adds pix2pixHD's lambda_feat * VGGLoss(fake, real) to the generator
        objective; None = --no_vgg_loss (the offline default: there is no ImageNet checkpoint to load)."""
        self.pipe, self.netD, self.vgg = pipe, netD, vgg
        self.batch_d = __import__("os").environ.get("NHVR_BATCH_D", "1") != "0"
        self.lam = dict(feat=lambda_feat, l2=lambda_l2, uv=lambda_uv, prob=lambda_prob, temp=lambda_temp)
        self.n_layers_D, self.num_D = n_layers_D, num_D
        # one flat bucket per network: its all-reduce starts when that network's backward ends and overlaps the rest
        self.buckets_G = {"netG": ParamBucket(pipe.netG.parameters(), lr, beta1, owners=[pipe.netG]),
                          "netBG": ParamBucket(pipe.netBG.parameters(), lr, beta1, owners=[pipe.netBG]),
                          "netTransG": ParamBucket(pipe.netTransG.parameters(), lr, beta1, owners=[pipe.netTransG]),
                          "tex": ParamBucket([pipe.atlas, pipe.bg], lr, beta1)}
        self.bucket_D = ParamBucket(netD.parameters(), lr, beta1, owners=[netD])
        self.distributed = distributed
        self._d_backwards = 0
        if distributed:
            for name in ("netG", "netBG", "netTransG"):
                getattr(pipe, name)._after_backward = self.buckets_G[name].all_reduce_async
                getattr(pipe, name)._mid_backward = self.buckets_G[name].all_reduce_tail_async
            netD._after_backward = self._d_backward_done

    def _d_backward_done(self) -> None:
        """The discriminator runs twice per scale under its own loss (fake detached, real): its bucket is complete after
        2 * num_D weight-gradient backwards."""
        self._d_backwards += 1
        if self._d_backwards == (1 if self.batch_d else 2) * self.num_D:
            self.bucket_D.all_reduce_async()

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
        if self.batch_d:
            # fake (detached) and real through the discriminator as ONE batch of 2N: InstanceNorm is per sample, so the values are
            # those of two separate passes, with half the launches and twice the tiles per launch
            both = D(torch.cat([pose1, pose1], 0), torch.cat([fake.detach(), real1], 0))
            n = fake.shape[0]
            pred_fake_d = [[t[:n] for t in s] for s in both]
            pred_real = [[t[n:] for t in s] for s in both]
        else:
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
        if self.vgg is not None:
            loss_G = loss_G + losses.vgg_diff(self.vgg, fake, real1, lam["feat"])
        # ---- generator side (the discriminator's weights are frozen under loss_G: its backward yields input grads only)
        for b in self.buckets_G.values():
            b.zero_grad()
        self.bucket_D.zero_grad()
        self._d_backwards = 0
        loss_G.backward()
        if self.distributed:
            self.buckets_G["tex"].all_reduce_async()
        # ---- discriminator (its all-reduce and the generator side's run under each other's backward / update)
        loss_D.backward()
        for b in self.buckets_G.values():
            b.adam_step()
        self.bucket_D.adam_step()
        pipe.ext_version = getattr(pipe, "ext_version", 0) + 1
        return {"loss_D": loss_D.detach(), "loss_G": loss_G.detach()}


def synthetic_train_batch(N: int, size: int, device, seed: int = 0) -> dict:
    """Synthetic sample of the end-to-end shapes (SURVEY §8d cfg 3): poses, real image U(-1,1), DensePose IUV, flow N(0, 2 px)."""
    pose, dp_i, dp_uv = synthetic_densepose(N, size, size, device, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    pose_prev = torch.roll(pose, shifts=3, dims=-1)
    image = torch.tanh(torch.nn.functional.interpolate(torch.randn(N, 3, 32, 32, generator=g), size=(size, size), mode="bilinear")).to(device)
    flow = (torch.randn(N, 2, size, size, generator=g) * 2.0).to(device)
    return {"pose_prev": pose_prev, "pose": pose, "image": image, "dp_i": dp_i, "dp_uv": dp_uv, "flow_inv": flow}
