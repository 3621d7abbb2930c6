"""Packs the reference's 100 bundled OpenPose fixtures [REF keypoints/frame000{00..99}_keypoints.json] into one
small array so that the GPU box (where /root/reference does not exist) can rasterise the real configs[0]
pose sequence:  tests/golden/keypoints_body25.npy  float32 [100, 25, 3] = (x, y, confidence) of the BODY_25 joints.
Run from the repo root in the build container:  python tests/golden/make_keypoints_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nhvr_b200 import pose  # noqa: E402

if __name__ == "__main__":
    kps = pose.read_sequence("/root/reference/keypoints")
    assert kps.shape == (100, 25, 3), kps.shape
    np.save(os.path.join(ROOT, "tests", "golden", "keypoints_body25.npy"), kps.astype(np.float32))
    print("wrote keypoints_body25.npy", kps.shape, float(kps[..., 2].min()))
