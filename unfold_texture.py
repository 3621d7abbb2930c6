"""`python3 unfold_texture.py $video_frame_dir $densepose_dir [--out texture.jpg] [--size 200]` [REF README.md:64]:
the initial texture.jpg (the 4 x 6 grid of 24 DensePose part textures, 800 x 1200 at part size 200) from video frames and
their DensePose IUV images.  The accumulation runs on the sm_100a scatter kernel (nhvr_texture_unfold, the texture
lookup's adjoint); reading the frame / IUV images is host code (cv2 or PIL).

DensePose IUV image convention (public DensePose `*_IUV.png`): channel 0 = part index I (0..24), channels 1, 2 = U, V * 255.
"""
import argparse
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch

from nhvr_b200 import capi, ops


def read_image(path):
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise IOError(path)
        return img[:, :, ::-1].copy()          # BGR -> RGB
    except ImportError:
        from PIL import Image
        return np.asarray(Image.open(path).convert("RGB"))


def atlas_to_grid(atlas: torch.Tensor) -> np.ndarray:
    """[24, 3, S, S] in [-1, 1] -> uint8 [4*S, 6*S, 3]: part k at row k // 6, column k % 6 (the DensePose atlas layout)."""
    P, C, S, _ = atlas.shape
    grid = atlas.reshape(4, 6, C, S, S).permute(0, 3, 1, 4, 2).reshape(4 * S, 6 * S, C)
    return ((grid.clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8).cpu().numpy()


def grid_to_atlas(img: np.ndarray, S: int) -> torch.Tensor:
    """inverse of atlas_to_grid: uint8 [4*S, 6*S, 3] -> [24, 3, S, S] in [-1, 1] (how --texture_path texture.jpg is loaded)."""
    t = torch.from_numpy(np.ascontiguousarray(img)).float() / 127.5 - 1.0
    return t.reshape(4, S, 6, S, 3).permute(0, 2, 4, 1, 3).reshape(24, 3, S, S).contiguous()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("video_frame_dir")
    ap.add_argument("densepose_dir")
    ap.add_argument("--out", default="texture.jpg")
    ap.add_argument("--size", type=int, default=200, help="part texture size [REF pre_train_tex.sh:19 --loadSize 200]")
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args(argv)
    capi.require_device()
    frames = sorted(glob.glob(os.path.join(a.video_frame_dir, "*.jpg")) + glob.glob(os.path.join(a.video_frame_dir, "*.png")))
    iuvs = sorted(glob.glob(os.path.join(a.densepose_dir, "*.png")))
    if not frames or len(frames) != len(iuvs):
        raise SystemExit("unfold_texture.py: need as many DensePose IUV pngs (%d) as frames (%d)" % (len(iuvs), len(frames)))
    unf = ops.TextureUnfolder(a.size, 3)
    for i in range(0, len(frames), a.batch):
        imgs = np.stack([read_image(p) for p in frames[i:i + a.batch]])
        iuv = np.stack([read_image(p) for p in iuvs[i:i + a.batch]])
        img = (torch.from_numpy(imgs).cuda().permute(0, 3, 1, 2).float() / 127.5 - 1.0)
        dp_i = torch.from_numpy(iuv[..., 0].astype(np.int32)).cuda()
        dp_uv = torch.from_numpy(iuv[..., 1:3].copy()).cuda().permute(0, 3, 1, 2).float() / 255.0
        unf.add(img, dp_i, dp_uv)
    grid = atlas_to_grid(unf.atlas())
    try:
        import cv2
        cv2.imwrite(a.out, grid[:, :, ::-1])
    except ImportError:
        from PIL import Image
        Image.fromarray(grid).save(a.out)
    print("[unfold_texture.py] %d frames -> %s (%d x %d)" % (len(frames), a.out, grid.shape[1], grid.shape[0]))


if __name__ == "__main__":
    main()
