"""OpenPose keypoint JSON -> pose-map tensors (the step just before the hot path).

Input format pinned by the fixture [REF keypoints/frame000{00..99}_keypoints.json]: OpenPose v1.2,
one person, ``pose_keypoints_2d`` = 25 BODY_25 joints x (x, y, conf) in a ~1024^2 frame
(SURVEY Appendix B).  The pose map is a 3-channel stick figure (``--input_nc 3`` [REF start.sh:24])
in [-1, 1]; limb list and colours are OpenPose's BODY_25 rendering convention (SPEC D1).
Host-side numpy: 25 joints per frame, not a hot loop.
"""
from __future__ import annotations

import glob
import json
import os
from typing import List, Optional, Tuple

import numpy as np

# BODY_25 limb pairs (OpenPose getPosePartPairs(BODY_25))
BODY25_PAIRS: List[Tuple[int, int]] = [
    (1, 8), (1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (8, 9), (9, 10), (10, 11), (8, 12), (12, 13), (13, 14),
    (1, 0), (0, 15), (15, 17), (0, 16), (16, 18), (14, 19), (19, 20), (14, 21), (11, 22), (22, 23), (11, 24)]


def _limb_colors(n: int) -> np.ndarray:
    """Distinct hues around the colour wheel, one per limb, uint8 RGB."""
    h = np.arange(n, dtype=np.float64) / n
    r = np.clip(np.abs(h * 6 - 3) - 1, 0, 1)
    g = np.clip(2 - np.abs(h * 6 - 2), 0, 1)
    b = np.clip(2 - np.abs(h * 6 - 4), 0, 1)
    return (np.stack([r, g, b], 1) * 255).astype(np.uint8)


LIMB_COLORS = _limb_colors(len(BODY25_PAIRS))


def read_keypoints(path: str) -> np.ndarray:
    """One OpenPose JSON -> [25, 3] float32 (x, y, conf); zeros if no person."""
    with open(path) as f:
        d = json.load(f)
    people = d.get("people", [])
    if not people:
        return np.zeros((25, 3), np.float32)
    k = np.asarray(people[0]["pose_keypoints_2d"], np.float32).reshape(-1, 3)
    out = np.zeros((25, 3), np.float32)
    out[:min(25, k.shape[0])] = k[:25]
    return out


def list_keypoint_files(pose_path: str) -> List[str]:
    return sorted(glob.glob(os.path.join(pose_path, "*_keypoints.json")))


def read_sequence(pose_path: str, limit: Optional[int] = None) -> np.ndarray:
    files = list_keypoint_files(pose_path)
    if limit is not None:
        files = files[:limit]
    if not files:
        raise FileNotFoundError("no *_keypoints.json under %s" % pose_path)
    return np.stack([read_keypoints(f) for f in files])


def align_to_target(src: np.ndarray, tgt: Optional[np.ndarray]) -> np.ndarray:
    """Scale + translate source skeletons to a target person's statistics [REF start.sh:10, README.md:36;
    data/data_prep/run_alignPose.sh:8-10 --target_spread/--source_spread/--calculate_scale_translation].
    Uses ankle height (joints 11, 14) and body height (nose 0 -> ankles) medians; identity if tgt is None."""
    if tgt is None or len(tgt) == 0:
        return src

    def stats(k):
        ank = (k[:, 11, 1] + k[:, 14, 1]) * 0.5
        height = ank - k[:, 0, 1]
        ok = (k[:, 0, 2] > 0) & (k[:, 11, 2] > 0) & (k[:, 14, 2] > 0)
        if not ok.any():
            return None
        cx = np.median(k[ok, 8, 0])
        return float(np.median(ank[ok])), float(np.median(height[ok])), float(cx)

    s, t = stats(src), stats(tgt)
    if s is None or t is None or s[1] <= 1e-3:
        return src
    scale = t[1] / s[1]
    out = src.copy()
    out[:, :, 0] = (src[:, :, 0] - s[2]) * scale + t[2]
    out[:, :, 1] = (src[:, :, 1] - s[0]) * scale + t[0]
    return out


def rasterize(kp: np.ndarray, size: int, src_size: float = 1024.0, thickness: float = 4.0,
              conf_thresh: float = 0.05) -> np.ndarray:
    """[25,3] keypoints in a src_size^2 frame -> [3, size, size] float32 stick figure in [-1, 1]."""
    s = size / float(src_size)
    img = np.zeros((size, size, 3), np.float32)
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float32)
    half = max(thickness * size / 512.0, 1.0) * 0.5
    for li, (a, b) in enumerate(BODY25_PAIRS):
        if kp[a, 2] < conf_thresh or kp[b, 2] < conf_thresh:
            continue
        ax, ay, bx, by = kp[a, 0] * s, kp[a, 1] * s, kp[b, 0] * s, kp[b, 1] * s
        x0, x1 = int(max(min(ax, bx) - half - 1, 0)), int(min(max(ax, bx) + half + 2, size))
        y0, y1 = int(max(min(ay, by) - half - 1, 0)), int(min(max(ay, by) + half + 2, size))
        if x0 >= x1 or y0 >= y1:
            continue
        px, py = xs[y0:y1, x0:x1], ys[y0:y1, x0:x1]
        dx, dy = bx - ax, by - ay
        L2 = dx * dx + dy * dy
        t = np.clip(((px - ax) * dx + (py - ay) * dy) / L2, 0, 1) if L2 > 1e-6 else np.zeros_like(px)
        d2 = (px - (ax + t * dx)) ** 2 + (py - (ay + t * dy)) ** 2
        m = d2 <= half * half
        img[y0:y1, x0:x1][m] = LIMB_COLORS[li].astype(np.float32)
    return (img.transpose(2, 0, 1) / 127.5 - 1.0).astype(np.float32)


def pose_maps(kps: np.ndarray, size: int, pose_nc: int = 3, src_size: float = 1024.0) -> np.ndarray:
    """[T,25,3] -> [T, pose_nc, size, size]; channels beyond 3 (Laplace, absent in the fixtures) are zero."""
    T = kps.shape[0]
    out = np.zeros((T, pose_nc, size, size), np.float32)
    for t in range(T):
        out[t, :3] = rasterize(kps[t], size, src_size)
    return out
