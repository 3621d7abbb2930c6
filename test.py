"""Inference entry point: `bash test_start/start.sh` runs unmodified against this file
[REF test_start/start.sh:6-28].  Thin driver: options -> keypoints -> pose maps -> RenderPipeline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch

from nhvr_b200 import capi
from nhvr_b200.options import TestOptions, pipeline_kwargs
from nhvr_b200.pipeline import RenderPipeline, shard_frames
from nhvr_b200 import pose as posemod
from nhvr_b200.checkpoint import load_pipeline


def save_frame(path, chw):
    img = ((chw.clamp(-1, 1) + 1) * 127.5).round().byte().permute(1, 2, 0).cpu().numpy()
    try:
        from PIL import Image
        Image.fromarray(img).save(path)
    except ImportError:
        np.save(os.path.splitext(path)[0] + ".npy", img)


def main(argv=None):
    opt = TestOptions().parse(argv)
    capi.require_device()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0])))
    pipe = RenderPipeline(**pipeline_kwargs(opt)).cuda()
    found = load_pipeline(pipe, os.path.join(opt.checkpoints_dir, opt.name), opt.which_epoch)
    if not found:
        print("[test.py] no checkpoint under %s/%s: using random-init weights" % (opt.checkpoints_dir, opt.name))
    kps = posemod.read_sequence(opt.pose_path, None if opt.how_many >= 10 ** 5 else opt.how_many)
    tgt = posemod.read_sequence(opt.pose_tgt_path) if opt.pose_tgt_path and os.path.isdir(opt.pose_tgt_path) else None
    kps = posemod.align_to_target(kps, tgt)
    T = kps.shape[0]
    os.makedirs(opt.results_dir, exist_ok=True)
    clips = shard_frames(T, world, rank, opt.clips_in_flight)
    L = max(b - a for a, b in clips)
    # lock-step clips need equal length: pad the shorter ones by repeating their last pose (frames dropped on save)
    maps = np.zeros((len(clips), L, opt.pose_nc, opt.loadSize, opt.loadSize), np.float32)
    for c, (a, b) in enumerate(clips):
        if b > a:
            m = posemod.pose_maps(kps[a:b], opt.loadSize, opt.pose_nc)
            maps[c, :b - a] = m
            maps[c, b - a:] = m[-1]
    poses = torch.from_numpy(maps).pin_memory()
    frames = pipe.render_clips(poses)
    torch.cuda.synchronize()
    for c, (a, b) in enumerate(clips):
        for t in range(b - a):
            save_frame(os.path.join(opt.results_dir, "frame%05d.png" % (a + t)), frames[c, t])
    print("[test.py] rank %d rendered %d frames -> %s" % (rank, sum(b - a for a, b in clips), opt.results_dir))


if __name__ == "__main__":
    main()
