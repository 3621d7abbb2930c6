"""ORACLE (test infrastructure) — PARITY UNPINNED.  Texture lookup + composite in fp32 torch / numpy.

Restates the reference's "--TexG part --use_mask_texture" texture stage [REF test_start/start.sh:13-14,18;
README.md:64 (texture.jpg = DensePose part atlas); pre_train_tex.sh:19 (part size 200)] and the
mask/background blend [REF README.md:15,52,60; start.sh:12].  The reference's own code for these is
absent; the formula is SPEC D5-D7:

    part   = argmax_k logits[k]                     k in 0..24, lowest index wins ties
    u      = clamp(0.5*U + 0.5, 0, 1)               "hard sigmoid": no transcendental in the index path
    fx     = u*(S-1); x0 = floor(fx); x1 = min(x0+1, S-1); wx = fx - x0        (same for v / y)
    sample = (1-wy)*((1-wx)*T[y0,x0] + wx*T[y0,x1]) + wy*((1-wx)*T[y1,x0] + wx*T[y1,x1])
    tex    = sum_{k=1..24} softmax(logits)[k] * sample_k     (/(1 - P0 + 1e-6) unless use_mask_texture)

``align_corners=True`` semantics (u*(S-1)) are those of grid_sample in the pinned torch 1.1.0
[REF requirment.txt:5].  The integer outputs (part, x0, y0) are the bit-exact contract.
"""
from __future__ import annotations

import numpy as np
import torch

N_PARTS = 24


def uv_activation(t: torch.Tensor) -> torch.Tensor:
    return torch.clamp(t * 0.5 + 0.5, 0.0, 1.0)


def texture_sample(uvp: torch.Tensor, atlas: torch.Tensor, use_mask_texture: bool = True):
    """uvp [N,73,H,W] fp32 (25 logits, 24 U, 24 V); atlas [24,Ctex,S,S] fp32.

    Returns tex [N,Ctex,H,W] fp32, part [N,H,W] uint8, texel [N,H,W,2] int16 (x0,y0 of argmax part).
    """
    assert uvp.dim() == 4 and uvp.shape[1] == 25 + 2 * N_PARTS
    N, _, H, W = uvp.shape
    P, Ctex, S, S2 = atlas.shape
    assert P == N_PARTS and S == S2
    uvp = uvp.float()
    atlas = atlas.float()
    logits = uvp[:, :25]
    part = torch.argmax(logits, dim=1)                       # first max wins (documented tie-break)
    prob = torch.softmax(logits, dim=1)
    u = uv_activation(uvp[:, 25:25 + N_PARTS])               # [N,24,H,W]
    v = uv_activation(uvp[:, 25 + N_PARTS:])
    fx = u * float(S - 1)
    fy = v * float(S - 1)
    x0f = torch.floor(fx)
    y0f = torch.floor(fy)
    x0 = x0f.long()
    y0 = y0f.long()
    x1 = torch.clamp(x0 + 1, max=S - 1)
    y1 = torch.clamp(y0 + 1, max=S - 1)
    wx = (fx - x0f).unsqueeze(2)                             # [N,24,1,H,W]
    wy = (fy - y0f).unsqueeze(2)

    flat = atlas.reshape(N_PARTS, Ctex, S * S)               # [24,C,S*S]

    def gather(yy, xx):
        idx = (yy * S + xx).reshape(N, N_PARTS, 1, H * W).expand(N, N_PARTS, Ctex, H * W)
        src = flat.unsqueeze(0).expand(N, N_PARTS, Ctex, S * S)
        return torch.gather(src, 3, idx).reshape(N, N_PARTS, Ctex, H, W)

    t00, t01, t10, t11 = gather(y0, x0), gather(y0, x1), gather(y1, x0), gather(y1, x1)
    sample = (1 - wy) * ((1 - wx) * t00 + wx * t01) + wy * ((1 - wx) * t10 + wx * t11)
    tex = (prob[:, 1:].unsqueeze(2) * sample).sum(dim=1)     # [N,C,H,W]
    if not use_mask_texture:
        tex = tex / (1.0 - prob[:, :1] + 1e-6)

    # integer contract
    k = torch.clamp(part - 1, min=0)
    x0p = torch.gather(x0, 1, k.unsqueeze(1)).squeeze(1)
    y0p = torch.gather(y0, 1, k.unsqueeze(1)).squeeze(1)
    fg = part > 0
    texel = torch.stack([torch.where(fg, x0p, torch.zeros_like(x0p)),
                         torch.where(fg, y0p, torch.zeros_like(y0p))], dim=-1).to(torch.int16)
    return tex, part.to(torch.uint8), texel


def texture_sample_numpy(uvp: np.ndarray, atlas: np.ndarray, use_mask_texture: bool = True):
    """Pure-numpy scalar-loop restatement for SMALL cases (pins the torch version above)."""
    N, _, H, W = uvp.shape
    _, Ctex, S, _ = atlas.shape
    uvp = uvp.astype(np.float32)
    atlas = atlas.astype(np.float32)
    tex = np.zeros((N, Ctex, H, W), np.float32)
    part = np.zeros((N, H, W), np.uint8)
    texel = np.zeros((N, H, W, 2), np.int16)
    half = np.float32(0.5)
    sm1 = np.float32(S - 1)
    for n in range(N):
        for y in range(H):
            for x in range(W):
                lg = uvp[n, :25, y, x]
                p = int(np.argmax(lg))
                e = np.exp(lg - lg.max())
                prob = e / e.sum()
                acc = np.zeros(Ctex, np.float64)
                for k in range(1, 25):
                    u = np.float32(min(max(np.float32(uvp[n, 24 + k, y, x] * half) + half, np.float32(0)), np.float32(1)))
                    v = np.float32(min(max(np.float32(uvp[n, 48 + k, y, x] * half) + half, np.float32(0)), np.float32(1)))
                    fx = np.float32(u * sm1)
                    fy = np.float32(v * sm1)
                    x0 = int(np.floor(fx)); y0 = int(np.floor(fy))
                    x1 = min(x0 + 1, S - 1); y1 = min(y0 + 1, S - 1)
                    wx = float(fx - np.float32(x0)); wy = float(fy - np.float32(y0))
                    T = atlas[k - 1]
                    s = (1 - wy) * ((1 - wx) * T[:, y0, x0] + wx * T[:, y0, x1]) + wy * ((1 - wx) * T[:, y1, x0] + wx * T[:, y1, x1])
                    acc += prob[k] * s
                    if k == p:
                        texel[n, y, x] = (x0, y0)
                if not use_mask_texture:
                    acc = acc / (1.0 - prob[0] + 1e-6)
                tex[n, :, y, x] = acc
                part[n, y, x] = p
    return tex, part, texel


def composite(fgm: torch.Tensor, bg: torch.Tensor) -> torch.Tensor:
    """out = m*fg + (1-m)*bg.  fgm [N,4,H,W] (RGB, mask), bg [3,H,W] or [N,3,H,W]  [REF README.md:15,52,60]."""
    fg, m = fgm[:, :3], fgm[:, 3:4]
    if bg.dim() == 3:
        bg = bg.unsqueeze(0)
    return m * fg + (1 - m) * bg


def unfold_texture(img: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor, S: int, min_weight: float = 0.0) -> torch.Tensor:
    """unfold_texture [REF README.md:64]: the initial part atlas from frames + DensePose IUV, as the weighted mean of the
    pixels splatted with the bilinear lookup's own corner weights (the adjoint of texture_sample's gather; the reference's
    script is absent - SPEC: same (u, v) -> texel rule as the lookup, align_corners).  img [N,C,H,W]; dp_i [N,H,W] in 0..24;
    dp_uv [N,2,H,W] in [0,1] -> atlas [24,C,S,S] (0 where no pixel landed)."""
    N, C, H, W = img.shape
    acc = torch.zeros(N_PARTS * S * S, C + 1, dtype=torch.float64)
    fg = (dp_i >= 1) & (dp_i <= N_PARTS)
    u = dp_uv[:, 0].float().clamp(0, 1)
    v = dp_uv[:, 1].float().clamp(0, 1)
    fx, fy = u * float(S - 1), v * float(S - 1)
    x0f, y0f = torch.floor(fx), torch.floor(fy)
    x0, y0 = x0f.long(), y0f.long()
    x1, y1 = torch.clamp(x0 + 1, max=S - 1), torch.clamp(y0 + 1, max=S - 1)
    wx, wy = (fx - x0f), (fy - y0f)
    vals = torch.cat([img.double().permute(0, 2, 3, 1), torch.ones(N, H, W, 1, dtype=torch.float64)], -1)   # [N,H,W,C+1]
    base = (dp_i.long() - 1).clamp(min=0) * S * S
    for yy, xx, w in ((y0, x0, (1 - wy) * (1 - wx)), (y0, x1, (1 - wy) * wx), (y1, x0, wy * (1 - wx)), (y1, x1, wy * wx)):
        at = (base + yy * S + xx)[fg]
        acc.index_add_(0, at, vals[fg] * w.double()[fg].unsqueeze(-1))
    wsum = acc[:, C:]
    atlas = torch.where(wsum > min_weight, acc[:, :C] / wsum.clamp(min=1e-30), torch.zeros_like(acc[:, :C]))
    return atlas.reshape(N_PARTS, S, S, C).permute(0, 3, 1, 2).float().contiguous()
