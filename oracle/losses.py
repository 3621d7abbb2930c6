"""ORACLE (test infrastructure) — PARITY UNPINNED.  Training losses in fp32 torch.

Restates the loss terms evidenced by the reference's training flags
[REF train_start/pretrain_start.sh:31-37: --lambda_L2 500 --lambda_UV 1000 --lambda_Prob 10
 --use_densepose_loss --lambda_Temp 500] plus the pix2pixHD GAN / feature-matching terms
[UPSTREAM GANLoss(use_lsgan=True), lambda_feat=10; SURVEY Appendix C].  Forms are SPEC D11 / D13.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .texture import uv_activation, N_PARTS


def gan_loss(pred_scales, target_is_real: bool) -> torch.Tensor:
    """LSGAN: MSE vs 1/0 on the LAST map of each scale, summed over scales [UPSTREAM GANLoss]."""
    loss = 0.0
    for pred in pred_scales:
        p = pred[-1] if isinstance(pred, (list, tuple)) else pred
        tgt = torch.ones_like(p) if target_is_real else torch.zeros_like(p)
        loss = loss + F.mse_loss(p, tgt)
    return loss


def feature_matching_loss(pred_fake, pred_real, n_layers_D: int = 3, num_D: int = 2, lambda_feat: float = 10.0):
    """sum_i sum_j (1/num_D) * (4/(n_layers_D+1)) * L1(fake_ij, real_ij.detach()) * lambda_feat [UPSTREAM]."""
    feat_w = 4.0 / (n_layers_D + 1)
    d_w = 1.0 / num_D
    loss = 0.0
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + d_w * feat_w * F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * lambda_feat
    return loss


def l2_loss(fake: torch.Tensor, real: torch.Tensor) -> torch.Tensor:
    """--lambda_L2 reconstruction term (mean squared error) [REF pretrain_start.sh:31]."""
    return F.mse_loss(fake, real)


def uv_loss(uvp: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor) -> torch.Tensor:
    """--lambda_UV: masked L1 between the predicted (u,v) of the ground-truth part and DensePose UV.

    uvp [N,73,H,W]; dp_i [N,H,W] long in 0..24 (0 = background); dp_uv [N,2,H,W] in [0,1].
    """
    N, _, H, W = uvp.shape
    k = torch.clamp(dp_i - 1, min=0).unsqueeze(1)
    u = torch.gather(uv_activation(uvp[:, 25:25 + N_PARTS]), 1, k).squeeze(1)
    v = torch.gather(uv_activation(uvp[:, 25 + N_PARTS:]), 1, k).squeeze(1)
    fg = (dp_i > 0).float()
    denom = fg.sum().clamp(min=1.0)
    return ((u - dp_uv[:, 0]).abs() * fg + (v - dp_uv[:, 1]).abs() * fg).sum() / denom


def prob_loss(uvp: torch.Tensor, dp_i: torch.Tensor) -> torch.Tensor:
    """--lambda_Prob: 25-way cross-entropy of the part logits vs DensePose I [REF pretrain_start.sh:33]."""
    return F.cross_entropy(uvp[:, :25], dp_i.long())


def flow_warp(img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """Bilinear backward warp: out(x) = img(x + flow(x)); flow [N,2,H,W] in pixels (dx, dy)."""
    N, _, H, W = img.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=img.device, dtype=img.dtype),
                            torch.arange(W, device=img.device, dtype=img.dtype), indexing="ij")
    gx = (xs.unsqueeze(0) + flow[:, 0]) / max(W - 1, 1) * 2 - 1
    gy = (ys.unsqueeze(0) + flow[:, 1]) / max(H - 1, 1) * 2 - 1
    grid = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)


def temporal_loss(out_t: torch.Tensor, out_prev: torch.Tensor, flow_inv: torch.Tensor) -> torch.Tensor:
    """--lambda_Temp: L1(out_t - warp(out_{t-1}, flow_inv_t)) [REF pretrain_start.sh:21-22,37; SPEC D11]."""
    return F.l1_loss(out_t, flow_warp(out_prev, flow_inv))


def vgg_loss(vgg, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """pix2pixHD ``VGGLoss``: sum_i w_i * L1(vgg(x)_i, vgg(y)_i.detach()), w = (1/32, 1/16, 1/8, 1/4, 1) [UPSTREAM models/networks.py]."""
    w = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)
    fx, fy = vgg(x), vgg(y)
    return sum(wi * torch.nn.functional.l1_loss(a, b.detach()) for wi, a, b in zip(w, fx, fy))
