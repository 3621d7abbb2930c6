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
