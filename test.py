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


def save_frame(path, hwc_u8):
    """One frame (uint8 HWC numpy) -> PNG; runs on the writer pool (cv2 / PIL release the GIL while encoding)."""
    try:
        import cv2
        cv2.imwrite(path, hwc_u8[:, :, ::-1], [cv2.IMWRITE_PNG_COMPRESSION, 1])
    except ImportError:
        try:
            from PIL import Image
            Image.fromarray(hwc_u8).save(path, compress_level=1)
        except ImportError:
            np.save(os.path.splitext(path)[0] + ".npy", hwc_u8)


def write_frames(frames, clips, results_dir, workers=None):
    """frames [C, L, 3, H, W] (pinned host fp32 in [-1, 1]) -> results_dir/frame%05d.png on a thread pool: at several
    hundred frames/s a serial PNG encoder is the bottleneck of the whole run."""
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or min(32, (os.cpu_count() or 4))

    def job(c, t, idx):
        img = ((frames[c, t].clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()
        save_frame(os.path.join(results_dir, "frame%05d.png" % idx), img)
    with ThreadPoolExecutor(workers) as pool:
        futs = [pool.submit(job, c, t, a + t) for c, (a, b) in enumerate(clips) for t in range(b - a)]
        for f in futs:
            f.result()


def main(argv=None):
    opt = TestOptions().parse(argv)
    capi.require_device()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0])))
    pipe = RenderPipeline(**pipeline_kwargs(opt), precision=opt.precision).cuda()
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
    # lock-step clips need equal length: pad the shorter ones by repeating their last pose (frames dropped on save).
    # Only the keypoints (300 bytes per frame) go to the GPU; the pose maps are rasterised there, inside every step.
    kpc = np.zeros((len(clips), L, 25, 3), np.float32)
    for c, (a, b) in enumerate(clips):
        if b > a:
            kpc[c, :b - a] = kps[a:b]
            kpc[c, b - a:] = kps[b - 1]
    frames = torch.empty(len(clips), L, 3, opt.loadSize, opt.loadSize).pin_memory()
    pipe.render_keypoints(torch.from_numpy(kpc), opt.loadSize, out=frames)
    torch.cuda.synchronize()
    capi.check_overflow()
    write_frames(frames, clips, opt.results_dir)
    print("[test.py] rank %d rendered %d frames -> %s" % (rank, sum(b - a for a, b in clips), opt.results_dir))


if __name__ == "__main__":
    main()
