"""`python3 graph_posenorm.py ...` exactly as data/data_prep/run_alignPose.sh:1-10 launches it: align a SOURCE person's
OpenPose keypoints to a TARGET person's frame (scale + translation from the ankle "spread" and the body height, the
pose-normalisation of "Everybody Dance Now" whose flag names the script carries), and write the aligned keypoint JSONs
(+ optional stick-figure images) to --results.  Pure host numpy: 25 joints per frame.

  --target_keypoints DIR   openpose_json of the target person          --target_shape H W C
  --source_keypoints DIR   openpose_json of the driving person         --source_shape H W C
  --source_frames DIR      (accepted; only used to name the outputs when present)
  --results DIR            output directory
  --target_spread LO HI / --source_spread LO HI   ankle-height range (pixels) of the frames used for the statistics
  --calculate_scale_translation                   derive scale / translation from the data (the only mode built)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from nhvr_b200 import pose as posemod


def spread_stats(k: np.ndarray, spread):
    """(close ankle y, far ankle y, median body height at the close / far end, median x of the mid hip) over the frames
    whose mean ankle height lies inside `spread` and whose nose + ankles were detected."""
    ank = 0.5 * (k[:, 11, 1] + k[:, 14, 1])
    ok = (k[:, 0, 2] > 0) & (k[:, 11, 2] > 0) & (k[:, 14, 2] > 0) & (ank >= spread[0]) & (ank <= spread[1])
    if not ok.any():
        return None
    ank, height, cx = ank[ok], (ank - k[:, 0, 1])[ok], k[ok, 8, 0]
    close, far = float(ank.max()), float(ank.min())
    tol = max(0.05 * (close - far), 1.0)
    h_close = float(np.median(height[ank >= close - tol]))
    h_far = float(np.median(height[ank <= far + tol]))
    return close, far, h_close, h_far, float(np.median(cx))


def align(src: np.ndarray, tgt: np.ndarray, source_spread, target_spread) -> np.ndarray:
    """Per-frame scale + translation: the ankle position interpolates between the target's far / close ankle lines, the
    scale between the far / close height ratios (EDN pose normalisation); x is re-centred on the target's mid-hip median."""
    s, t = spread_stats(src, source_spread), spread_stats(tgt, target_spread)
    if s is None or t is None:
        return posemod.align_to_target(src, tgt)
    s_close, s_far, sh_close, sh_far, s_cx = s
    t_close, t_far, th_close, th_far, t_cx = t
    out = src.copy()
    ank = 0.5 * (src[:, 11, 1] + src[:, 14, 1])
    a = np.clip((ank - s_far) / max(s_close - s_far, 1e-6), 0.0, 1.0)          # 0 = far, 1 = close
    scale = (1 - a) * (th_far / max(sh_far, 1e-6)) + a * (th_close / max(sh_close, 1e-6))
    new_ank = t_far + a * (t_close - t_far)
    out[:, :, 0] = (src[:, :, 0] - s_cx) * scale[:, None] + t_cx
    out[:, :, 1] = (src[:, :, 1] - ank[:, None]) * scale[:, None] + new_ank[:, None]
    out[:, :, 2] = src[:, :, 2]
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--target_keypoints", required=True)
    ap.add_argument("--source_keypoints", required=True)
    ap.add_argument("--target_shape", type=int, nargs=3, default=[1024, 1024, 3])
    ap.add_argument("--source_shape", type=int, nargs=3, default=[1024, 1024, 3])
    ap.add_argument("--source_frames", default="")
    ap.add_argument("--results", required=True)
    ap.add_argument("--target_spread", type=float, nargs=2, default=[400, 800])
    ap.add_argument("--source_spread", type=float, nargs=2, default=[400, 800])
    ap.add_argument("--calculate_scale_translation", action="store_true")
    ap.add_argument("--draw", action="store_true", help="also write the aligned stick-figure images")
    a = ap.parse_args(argv)
    src_files = posemod.list_keypoint_files(a.source_keypoints)
    src, tgt = posemod.read_sequence(a.source_keypoints), posemod.read_sequence(a.target_keypoints)
    # bring the source into the target's pixel frame first (different capture resolutions)
    src = src.copy()
    src[:, :, 0] *= a.target_shape[1] / float(a.source_shape[1])
    src[:, :, 1] *= a.target_shape[0] / float(a.source_shape[0])
    sc = a.target_shape[0] / float(a.source_shape[0])
    out = align(src, tgt, [v * sc for v in a.source_spread], a.target_spread) if a.calculate_scale_translation else src
    os.makedirs(a.results, exist_ok=True)
    for f, k in zip(src_files, out):
        d = json.load(open(f))
        if d.get("people"):
            d["people"][0]["pose_keypoints_2d"] = [float(v) for v in k.reshape(-1)]
        json.dump(d, open(os.path.join(a.results, os.path.basename(f)), "w"))
        if a.draw:
            img = ((posemod.rasterize(k, a.target_shape[0], float(a.target_shape[0])).transpose(1, 2, 0) + 1) * 127.5).astype(np.uint8)
            np.save(os.path.join(a.results, os.path.basename(f).replace("_keypoints.json", "_pose.npy")), img)
    print("[graph_posenorm.py] %d source frames aligned to %d target frames -> %s" % (len(out), len(tgt), a.results))
    return out


if __name__ == "__main__":
    main()
