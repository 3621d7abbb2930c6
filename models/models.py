"""pix2pixHD's ``create_model(opt)`` entry: returns the rendering pipeline built from the parsed flags."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nhvr_b200.options import pipeline_kwargs  # noqa: E402
from nhvr_b200.pipeline import RenderPipeline  # noqa: E402


def create_model(opt):
    return RenderPipeline(**pipeline_kwargs(opt)).cuda()
