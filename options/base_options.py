"""Reference-compatible import path (pix2pixHD ``options/`` layout); see nhvr_b200/options.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nhvr_b200.options import BaseOptions, TrainOptions, TestOptions  # noqa: F401,E402
