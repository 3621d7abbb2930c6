"""Training losses of the reference on the sm_100a reduction kernels (forward values; fp32 in, fp64 accumulate).

Terms and weights are the reference's flags [REF train_start/pretrain_start.sh:31-37: --lambda_L2 500,
--lambda_UV 1000, --lambda_Prob 10, --use_densepose_loss, --lambda_Temp 500] plus pix2pixHD's LSGAN and
feature-matching terms [SURVEY Appendix C].  Same formulas as oracle/losses.py (SPEC D11, D13); parity bound
1e-3 relative (BASELINE.json north_star).  Each function returns a 0-dim float64 CUDA tensor.
"""
from __future__ import annotations

from typing import Sequence

import torch

from .capi import check, load, stream_ptr


def _acc(dev, n=1):
    return torch.zeros(n, dtype=torch.float64, device=dev)


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "nhvr_b200 losses take CUDA tensors only (no CPU fallback)"
    return t.detach().contiguous().float()


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    acc = _acc(a.device)
    check(load().nhvr_loss_sum_sq_diff(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), stream_ptr()), "nhvr_loss_sum_sq_diff")
    return acc[0] / a.numel()


def l1(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    acc = _acc(a.device)
    check(load().nhvr_loss_sum_abs_diff(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), stream_ptr()), "nhvr_loss_sum_abs_diff")
    return acc[0] / a.numel()


def gan_loss(pred_scales: Sequence, target_is_real: bool) -> torch.Tensor:
    """LSGAN: MSE vs 1/0 on the last map of each scale, summed over scales."""
    total = None
    for pred in pred_scales:
        p = _f32(pred[-1] if isinstance(pred, (list, tuple)) else pred)
        acc = _acc(p.device)
        check(load().nhvr_loss_sum_sq_const(p.data_ptr(), 1.0 if target_is_real else 0.0, p.numel(), acc.data_ptr(), stream_ptr()),
              "nhvr_loss_sum_sq_const")
        term = acc[0] / p.numel()
        total = term if total is None else total + term
    return total


def feature_matching_loss(pred_fake, pred_real, n_layers_D: int = 3, num_D: int = 2, lambda_feat: float = 10.0) -> torch.Tensor:
    feat_w, d_w = 4.0 / (n_layers_D + 1), 1.0 / num_D
    total = None
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            term = l1(pred_fake[i][j], pred_real[i][j]) * (d_w * feat_w * lambda_feat)
            total = term if total is None else total + term
    return total


def uv_prob_losses(uvp: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor):
    """(uv_loss, prob_loss): masked L1 of the ground-truth part's (u,v) vs DensePose UV; 25-way part cross-entropy."""
    uvp, dp_uv = _f32(uvp), _f32(dp_uv)
    dp = dp_i.detach().contiguous().to(torch.int32)
    N, _, H, W = uvp.shape
    acc = _acc(uvp.device, 3)
    check(load().nhvr_loss_uv_prob(uvp.data_ptr(), dp.data_ptr(), dp_uv.data_ptr(), N, H, W, acc.data_ptr(), stream_ptr()),
          "nhvr_loss_uv_prob")
    return acc[0] / torch.clamp(acc[1], min=1.0), acc[2] / (N * H * W)


def temporal_loss(out_t: torch.Tensor, out_prev: torch.Tensor, flow_inv: torch.Tensor) -> torch.Tensor:
    cur, prev, fl = _f32(out_t), _f32(out_prev), _f32(flow_inv)
    N, C, H, W = cur.shape
    acc = _acc(cur.device)
    check(load().nhvr_loss_temporal(cur.data_ptr(), prev.data_ptr(), fl.data_ptr(), N, C, H, W, acc.data_ptr(), stream_ptr()),
          "nhvr_loss_temporal")
    return acc[0] / cur.numel()


class _UVProbLoss(torch.autograd.Function):
    """lambda_UV * uv_loss + lambda_Prob * prob_loss with its gradient w.r.t. the UV-generator output, both on
    the sm_100a reduction kernels (the UV-generator pre-train objective, REF pretrainTrans.sh; pretrain_start.sh:32-33)."""

    @staticmethod
    def forward(ctx, uvp, dp_i, dp_uv, w_uv, w_prob):
        u, d = _f32(uvp), _f32(dp_uv)
        dp = dp_i.detach().contiguous().to(torch.int32)
        N, _, H, W = u.shape
        acc = _acc(u.device, 3)
        check(load().nhvr_loss_uv_prob(u.data_ptr(), dp.data_ptr(), d.data_ptr(), N, H, W, acc.data_ptr(), stream_ptr()),
              "nhvr_loss_uv_prob")
        ctx.saved = (u, dp, d, acc, float(w_uv), float(w_prob))
        loss = w_uv * acc[0] / torch.clamp(acc[1], min=1.0) + w_prob * acc[2] / (N * H * W)
        return loss.float()

    @staticmethod
    def backward(ctx, grad_out):
        u, dp, d, acc, w_uv, w_prob = ctx.saved
        N, _, H, W = u.shape
        grad = torch.empty_like(u)
        gs = grad_out.detach().reshape(1).float().contiguous()
        check(load().nhvr_loss_uv_prob_bwd(u.data_ptr(), dp.data_ptr(), d.data_ptr(), N, H, W, acc.data_ptr(), w_uv, w_prob,
                                           gs.data_ptr(), grad.data_ptr(), stream_ptr()), "nhvr_loss_uv_prob_bwd")
        return grad, None, None, None, None


def uv_prob_objective(uvp, dp_i, dp_uv, lambda_uv: float = 1000.0, lambda_prob: float = 10.0) -> torch.Tensor:
    """Differentiable  lambda_UV * uv_loss + lambda_Prob * prob_loss  (0-dim fp32 CUDA tensor)."""
    return _UVProbLoss.apply(uvp, dp_i, dp_uv, lambda_uv, lambda_prob)


class _PairLoss(torch.autograd.Function):
    """coef * mean(f(a, b)) with gradient w.r.t. a (b is a target / detached): mode 0 MSE, 1 L1, 2 MSE vs constant."""

    @staticmethod
    def forward(ctx, a, b, mode, target, coef):
        x = _f32(a)
        y = _f32(b) if b is not None else None
        acc = _acc(x.device)
        lib = load()
        if mode == 0:
            check(lib.nhvr_loss_sum_sq_diff(x.data_ptr(), y.data_ptr(), x.numel(), acc.data_ptr(), stream_ptr()), "nhvr_loss_sum_sq_diff")
        elif mode == 1:
            check(lib.nhvr_loss_sum_abs_diff(x.data_ptr(), y.data_ptr(), x.numel(), acc.data_ptr(), stream_ptr()), "nhvr_loss_sum_abs_diff")
        else:
            check(lib.nhvr_loss_sum_sq_const(x.data_ptr(), float(target), x.numel(), acc.data_ptr(), stream_ptr()), "nhvr_loss_sum_sq_const")
        ctx.saved = (x, y, int(mode), float(target), float(coef))
        return (coef * acc[0] / x.numel()).float()

    @staticmethod
    def backward(ctx, gout):
        x, y, mode, target, coef = ctx.saved
        g = torch.empty_like(x)
        gs = gout.detach().reshape(1).float().contiguous()
        check(load().nhvr_loss_pair_bwd(x.data_ptr(), y.data_ptr() if y is not None else None, x.numel(), mode, target, coef,
                                        gs.data_ptr(), 0, g.data_ptr(), stream_ptr()), "nhvr_loss_pair_bwd")
        return g, None, None, None, None


def mse_diff(a, b, coef: float = 1.0):
    return _PairLoss.apply(a, b.detach(), 0, 0.0, coef)


def l1_diff(a, b, coef: float = 1.0):
    return _PairLoss.apply(a, b.detach(), 1, 0.0, coef)


def lsgan_diff(pred_scales, target_is_real: bool, coef: float = 1.0):
    total = None
    for pred in pred_scales:
        p = pred[-1] if isinstance(pred, (list, tuple)) else pred
        term = _PairLoss.apply(p, None, 2, 1.0 if target_is_real else 0.0, coef)
        total = term if total is None else total + term
    return total


def feature_matching_diff(pred_fake, pred_real, n_layers_D: int = 3, num_D: int = 2, lambda_feat: float = 10.0):
    feat_w, d_w = 4.0 / (n_layers_D + 1), 1.0 / num_D
    total = None
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            term = l1_diff(pred_fake[i][j], pred_real[i][j], d_w * feat_w * lambda_feat)
            total = term if total is None else total + term
    return total


class _TemporalLoss(torch.autograd.Function):
    """coef * L1(out_t - warp(out_{t-1}.detach(), flow_inv_t))  [REF pretrain_start.sh:21-22,37 --lambda_Temp; SPEC D11]."""

    @staticmethod
    def forward(ctx, cur, prev, flow, coef):
        c, p, f = _f32(cur), _f32(prev), _f32(flow)
        N, Cc, H, W = c.shape
        acc = _acc(c.device)
        check(load().nhvr_loss_temporal(c.data_ptr(), p.data_ptr(), f.data_ptr(), N, Cc, H, W, acc.data_ptr(), stream_ptr()),
              "nhvr_loss_temporal")
        ctx.saved = (c, p, f, float(coef))
        return (coef * acc[0] / c.numel()).float()

    @staticmethod
    def backward(ctx, gout):
        c, p, f, coef = ctx.saved
        N, Cc, H, W = c.shape
        g = torch.empty_like(c)
        gs = gout.detach().reshape(1).float().contiguous()
        check(load().nhvr_loss_temporal_bwd(c.data_ptr(), p.data_ptr(), f.data_ptr(), N, Cc, H, W, coef, gs.data_ptr(), g.data_ptr(),
                                            stream_ptr()), "nhvr_loss_temporal_bwd")
        return g, None, None, None


def temporal_diff(out_t, out_prev, flow_inv, coef: float = 1.0):
    return _TemporalLoss.apply(out_t, out_prev.detach(), flow_inv, coef)


VGG_WEIGHTS = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)


def vgg_diff(vgg, x: torch.Tensor, y: torch.Tensor, coef: float = 1.0):
    """pix2pixHD ``VGGLoss``: sum_i w_i * L1(vgg(x)_i, vgg(y)_i.detach()), w = (1/32, 1/16, 1/8, 1/4, 1), gradient to x only
    (networks.Vgg19B200; on unless --no_vgg_loss in the reference's training options)."""
    fx = vgg(x)
    with torch.no_grad():
        fy = vgg(y.detach())
    total = None
    for w, a, b in zip(VGG_WEIGHTS, fx, fy):
        t = l1_diff(a, b.detach(), coef * w)
        total = t if total is None else total + t
    return total
