"""ORACLE (test infrastructure) — PARITY UNPINNED.  The whole rendering path in fp32 torch.

Call stack restated from SURVEY.md §3.1 (inferred from REF test_start/start.sh:6-28):
    uvp  = TransG(pose)                      UV generator             [REF pretrainTrans.sh:13]
    tex  = sample(atlas, uvp)                "--TexG part"            [REF start.sh:13-14,18]
    fgm  = G(cat(tex, pose, prev))           temporal generator       [REF start.sh:7,15-17]
    bg'  = BG(bg)                            background refinement    [REF start.sh:12,20-21]
    out  = m*fg + (1-m)*bg'                  composite                [REF README.md:15,52,60]
    prev <- out                              (zeros at clip start; SPEC D8)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .networks import define_G, UV_CHANNELS, N_PARTS
from .texture import texture_sample, composite


class RenderModel(nn.Module):
    def __init__(self, pose_nc: int = 3, tex_nc: int = 3, size: int = 512, atlas_size: int = 200,
                 ngf_global: int = 48, n_downsample_global: int = 2, n_blocks_global: int = 10,
                 ngf_translate: int = 64, n_downsample_translate: int = 2, n_blocks_translate: int = 5,
                 ngf_bg: int = 48, n_downsample_bg: int = 2, n_blocks_bg: int = 2, use_mask_texture: bool = True):
        super().__init__()
        self.pose_nc, self.tex_nc, self.size = pose_nc, tex_nc, size
        self.use_mask_texture = use_mask_texture
        self.netTransG = define_G(pose_nc, UV_CHANNELS, ngf_translate, "translate", n_downsample_translate,
                                  n_blocks_translate)
        self.netG = define_G(tex_nc + pose_nc + 3, 4, ngf_global, "temporal", n_downsample_global, n_blocks_global)
        self.netBG = define_G(3, 3, ngf_bg, "bg", n_downsample_bg, n_blocks_bg)
        self.atlas = nn.Parameter(torch.empty(N_PARTS, tex_nc, atlas_size, atlas_size).uniform_(-1, 1))
        self.bg = nn.Parameter(torch.empty(3, size, size).uniform_(-1, 1))

    def refine_bg(self) -> torch.Tensor:
        return self.netBG(self.bg.unsqueeze(0))[0]

    def render_frame(self, pose: torch.Tensor, prev: torch.Tensor, bg_refined: torch.Tensor):
        uvp = self.netTransG(pose)
        tex, part, texel = texture_sample(uvp, self.atlas, self.use_mask_texture)
        fgm = self.netG(torch.cat([tex, pose, prev], dim=1))
        out = composite(fgm, bg_refined)
        return {"out": out, "fgm": fgm, "tex": tex, "uvp": uvp, "part": part, "texel": texel}

    @torch.no_grad()
    def render_clip(self, poses: torch.Tensor) -> torch.Tensor:
        """poses [T, pose_nc, H, W] -> frames [T, 3, H, W]; previous-frame state starts at zeros."""
        bg_refined = self.refine_bg()
        prev = torch.zeros(1, 3, poses.shape[-2], poses.shape[-1], dtype=poses.dtype, device=poses.device)
        frames = []
        for t in range(poses.shape[0]):
            r = self.render_frame(poses[t:t + 1], prev, bg_refined)
            prev = r["out"]
            frames.append(prev)
        return torch.cat(frames, dim=0)
