"""bench.py - frames/s rendered @512^2 through the hot path (BASELINE.json `metric`), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--frames_per_step F] [--precision strict] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: B independent lock-step clips (default 8) each
advance F frames (default 16) at 512 x 512 on one GPU - keypoints -> pose-map rasteriser -> UV generator -> texture
lookup -> temporal generator (previous-frame conditioned) -> composite with the refined background: B*F frames per step
per GPU.  This is BASELINE.json configs[3]'s per-GPU shard (clip-sharded long-sequence rendering with background
compositing) at the reference's real flags (test_start/start.sh: 6 pose channels).  Weights are random-init, the driving
skeleton is a random walk of the bundled frame-0 skeleton (stated in `data`).

value   : frames/s with the step's keypoints already resident in HBM (CUDA events around each step; every step touches
          B*F*(pose 6.3 MB + frame 3.1 MB) of fresh inputs/outputs and ~2 GB of activations - far beyond L2), max over ranks.
e2e     : the same metric through the public API RenderPipeline.render_keypoints with pinned HOST keypoints in and pinned
          HOST frames out (the device->host copy of every frame inside the timed region).
precision: the headline is the "strict" preset (split precision everywhere: the mode that meets north_star's 2e-2 / 45 dB
          on the reference's real configuration, tests/test_gpu_split3.py); `modes` reports "strict2", "balanced" and "fast" labelled.
roofline: the tcgen05 conv kernel (dominant): ALGORITHMIC conv FLOPs of a frame step / summed conv-launch durations
          (CUDA events around every launch on the launching stream, separate un-graphed pass), per layer class too.
cpu_baseline / --impl reference: the fp32 torch oracle (the only runnable statement of the reference's path - its source
          is absent) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

SIZE = 512
PIPE_KW = dict(pose_nc=6, tex_nc=3, size=SIZE, atlas_size=200, ngf_global=48, n_downsample_global=2,
               n_blocks_global=10, ngf_translate=64, n_downsample_translate=2, n_blocks_translate=5, ngf_bg=48,
               n_downsample_bg=2, n_blocks_bg=2, use_mask_texture=True)
PARITY_NOTE = {"strict": "frame parity 1.5e-3 max-abs / 93 dB (texture-like atlas), 4.3e-2 / 64 dB (U(-1,1) atlas; fp32 oracle's own noise 3.4e-3..6.9e-3)",
               "strict2": "frame parity 5.5e-3..7.9e-3 max-abs / 74-75 dB (texture-like atlas; temporal generator with 16-bit weights: 2 MMAs per product)",
               "balanced": "frame parity ~1.5e-2 max-abs / 67 dB (texture-like atlas)",
               "fast": "frame parity ~0.2 max-abs / 51 dB (texture-like atlas): UV error 3e-2"}


def bundled_keypoints():
    return np.load(os.path.join(ROOT, "tests", "golden", "keypoints_body25.npy"))


def synthetic_keypoints(B, T, seed=0):
    """[B, T, 25, 3]: random walk of the bundled frame-0 skeleton, ~5.7 px/frame (SURVEY 8d cfg 4), confidences kept."""
    g = np.random.default_rng(seed)
    k0 = bundled_keypoints()[0]
    walk = np.cumsum(g.normal(0.0, 4.0, size=(B, T, 1, 2)), axis=1) + np.cumsum(g.normal(0.0, 1.5, size=(B, T, 25, 2)), axis=1)
    walk = np.clip(walk, -150, 150)
    out = np.broadcast_to(k0, (B, T, 25, 3)).copy()
    out[..., :2] += walk.astype(np.float32)
    return torch.from_numpy(out.astype(np.float32))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                if len(f) > 6:
                    pw.append(float(f[6]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(clips, fps_step, precision):
    return {"workload": "configs[3] per-GPU shard: temporal clip rendering 512x512 with background compositing at "
                        "test_start/start.sh's flags (keypoints -> pose rasteriser (6 pose channels) -> UV generator ngf64/5 blocks "
                        "-> 24-part texture lookup 200^2 atlas -> temporal generator ngf48/2 down/10 blocks -> composite)",
            "clips_in_flight_per_gpu": clips, "frames_per_clip_per_step": fps_step, "frames_per_step_per_gpu": clips * fps_step,
            "resolution": SIZE, "precision": precision,
            "l2": "inputs larger than L2: every step reads/writes %d MB of fresh pose maps + frames and ~2 GB of activations per frame step"
                  % (clips * fps_step * (PIPE_KW["pose_nc"] + 3) * SIZE * SIZE * 4 // (1 << 20)),
            "parallelism": "clip-sharded, no collective"}


# ----------------------------------------------------------------------------------------------
def oracle_poses(T):
    from nhvr_b200 import pose as posemod
    kps = synthetic_keypoints(1, T)[0].numpy()
    return torch.from_numpy(posemod.pose_maps(kps, SIZE, PIPE_KW["pose_nc"]))


def cpu_oracle_fps(max_seconds=15.0, max_frames=8):
    """fp32 oracle on the host cores: frames/s over a bounded sample of the same per-frame path, batch 1."""
    from oracle.pipeline import RenderModel
    torch.manual_seed(0)
    model = RenderModel(**PIPE_KW).eval()
    poses = oracle_poses(4)
    with torch.no_grad():
        bg = model.refine_bg()
        prev = model.render_frame(poses[0:1], torch.zeros(1, 3, SIZE, SIZE), bg)["out"]      # warm-up frame
        n, t0 = 0, time.perf_counter()
        while True:
            prev = model.render_frame(poses[(n + 1) % 4:(n + 1) % 4 + 1], prev, bg)["out"]
            n += 1
            el = time.perf_counter() - t0
            if el > max_seconds or n >= max_frames:
                break
    return n / el, n, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the oracle port on the host cores (the reference's own source is absent).  A step is a bounded
    sample of the GPU arm's step: one frame of one clip (the per-frame path is identical).  Rank 0 only; the CPU arm does
    not scale with the GPU count, so it always reports n_gpus = 1."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle.pipeline import RenderModel
    torch.manual_seed(0)
    model = RenderModel(**PIPE_KW).eval()
    poses = oracle_poses(4)
    steps, warm = max(1, min(args.steps, 8)), 1
    with torch.no_grad():
        bg = model.refine_bg()
        prev = torch.zeros(1, 3, SIZE, SIZE)
        for i in range(warm):
            prev = model.render_frame(poses[i % 4:i % 4 + 1], prev, bg)["out"]
        t0 = time.perf_counter()
        for i in range(steps):
            prev = model.render_frame(poses[i % 4:i % 4 + 1], prev, bg)["out"]
        el = time.perf_counter() - t0
    fps = steps / el
    sample = ("%d step(s) of ONE frame of one clip each (a bounded sample of the GPU arm's %d-frame step; same per-frame path), "
              "fp32 torch oracle, %d threads" % (steps, args.clips * args.frames_per_step, cores))
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec rendered @512x512", "value": fps, "unit": "frames/s", "n_gpus": 1,
        "n_gpus_requested": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * el / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic keypoint random walk, random-init weights",
        "config": workload_config(args.clips, args.frames_per_step, "fp32 oracle"),
        "note": "CPU arm: does not scale with --gpus (rank 0's host cores only)",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------
def time_steps(pipe, kps_dev, B, F, K, Wm, barrier, sampler=None):
    """Device-resident timing of K steps (each F frame steps of B clips) after Wm warm-up steps; returns total ms, launches."""
    from nhvr_b200 import capi, ops
    step = pipe.step_graph(B, SIZE, SIZE, use_graph=True)
    step.reset()
    t = 0
    for _ in range(Wm * F):
        ops.pose_rasterize(kps_dev[:, t].contiguous(), SIZE, pipe.pose_nc, out=step.pose); step.run(); t += 1
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    if sampler is not None:
        sampler.start()
    n0 = capi.launch_count()
    for i in range(K):
        ev[i][0].record()
        for _ in range(F):
            ops.pose_rasterize(kps_dev[:, t].contiguous(), SIZE, pipe.pose_nc, out=step.pose); step.run(); t += 1
        ev[i][1].record()
    barrier()
    launches = (capi.launch_count() - n0) + step.launches_per_step * K * F      # graph replays are not counted by the library
    return sum(a.elapsed_time(b) for a, b in ev), launches, step


def conv_class(label):
    name = label.split(":", 1)[1]
    kind = name.split("_")[0].rstrip("0123456789x")           # stem / down / res / up / head
    cout = int(name.rsplit("-", 1)[1])
    net = "uv" if cout in (64, 128, 256, 73) else "g"
    return net + "_" + kind


def roofline_pass(pipe, kps_dev, B, pk, pk_src, timed_seconds):
    """Per-launch CUDA events over 3 un-graphed frame steps: conv aggregate, per layer class, and the memory kernels."""
    from nhvr_b200 import ops
    eager = pipe.step_graph(B, SIZE, SIZE, use_graph=False)
    eager.reset()
    for t in range(2):
        ops.pose_rasterize(kps_dev[:, t].contiguous(), SIZE, pipe.pose_nc, out=eager.pose); eager.run()
    torch.cuda.synchronize()
    ops.PROFILE = []
    REP = 3
    for t in range(REP):
        ops.pose_rasterize(kps_dev[:, t].contiguous(), SIZE, pipe.pose_nc, out=eager.pose); eager.run()
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    agg, cls = {}, {}
    for kind, work, a, b in recs:
        dt = a.elapsed_time(b) * 1e-3
        base = kind.split(":")[0]
        d = agg.setdefault("conv" if base.startswith("conv") else base, [0.0, 0.0, 0])
        d[0] += work; d[1] += dt; d[2] += 1
        if base.startswith("conv"):
            c = cls.setdefault(conv_class(kind), [0.0, 0.0, 0, 3.0 if base == "conv3" else 1.0])
            c[0] += work; c[1] += dt; c[2] += 1
    total = sum(v[1] for v in agg.values())
    conv = agg["conv"]
    ach = conv[0] / conv[1] / 1e12
    long_run = timed_seconds >= 2.0
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) if long_run else pk["bf16_tflops"]
    mma_mult = sum(c[0] * c[3] for c in cls.values()) / max(conv[0], 1.0)
    roof = {"kernel": "conv_shiftgemm_kernel (tcgen05)", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "peak_source": pk_src + (", sustained (timed region %.1f s)" % timed_seconds if long_run else ", burst (timed region %.2f s)" % timed_seconds),
            "achieved_mma_executed": ach * mma_mult, "frac_mma_executed": ach * mma_mult / peak,
            "note": "achieved = ALGORITHMIC conv FLOPs (2*k*k*Cin*Cout*Ho*Wo, un-padded) / conv time; split-precision layers execute 3 MMAs per "
                    "algorithmic one (achieved_mma_executed)",
            "traffic": None, "launches_per_frame_step": conv[2] // REP, "flops_per_frame_step": conv[0] / REP,
            "share_of_step": conv[1] / total,
            "per_class": {k: {"tflops": v[0] / v[1] / 1e12, "tflops_mma_executed": v[3] * v[0] / v[1] / 1e12, "ms_per_frame_step": 1e3 * v[1] / REP,
                              "launches": v[2] // REP} for k, v in sorted(cls.items())}}
    tp = os.path.join(ROOT, "profiles", "r02_conv_dram_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        roof["traffic"], roof["traffic_source"] = tj.get("mean_dram_bytes_per_launch"), tj.get("source")
    kernels = {}
    for kind in ("sampler", "composite", "in_apply", "pack"):
        if kind in agg:
            w, s, n = agg[kind]
            kernels[kind] = {"bound": "hbm", "achieved": w / s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": w / s / 1e9 / pk["hbm_gbs"],
                             "launches_per_frame_step": n // REP, "ms_per_frame_step": 1e3 * s / REP}
    # the RGB + mask head is HBM-bound by arithmetic intensity (AI 138 F/B): report it against HBM as well
    for k, v in cls.items():
        if k == "g_head":
            bytes_io = B * SIZE * SIZE * (48 * 2 * (2 if v[3] > 1 else 1) + 4 * 4)
            kernels["rgb_head"] = {"bound": "hbm", "achieved": bytes_io * v[2] / v[1] / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                   "frac": bytes_io * v[2] / v[1] / 1e9 / pk["hbm_gbs"], "ms_per_frame_step": 1e3 * v[1] / REP}
    return roof, kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=8, help="independent clips advanced in lock-step per GPU")
    ap.add_argument("--frames_per_step", type=int, default=16, help="frames each clip advances per step")
    ap.add_argument("--precision", type=str, default="strict", choices=["strict", "strict2", "balanced", "fast"])
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_train", action="store_true", help="skip the training legs")
    ap.add_argument("--no_legs", action="store_true", help="skip the configs[0] / configs[4] / other-precision legs")
    ap.add_argument("--train_batch", type=int, default=16)
    ap.add_argument("--train_e2e_batch", type=int, default=8, help="configs[2] batch per GPU (0 = skip that leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    from nhvr_b200 import capi, ops
    from nhvr_b200.pipeline import RenderPipeline
    capi.require_device()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly ONE JSON line: keep NCCL's banner out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(ms):
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    B, F, K, Wm = args.clips, args.frames_per_step, args.steps, args.warmup
    torch.manual_seed(0)
    pipe = RenderPipeline(**PIPE_KW, precision=args.precision).to(dev)
    T = (K + Wm) * F
    kps_host = synthetic_keypoints(B, T, seed=rank).pin_memory()
    kps_dev = kps_host.to(dev)

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local)
    dev_ms, launches, step = time_steps(pipe, kps_dev, B, F, K, Wm, barrier, sampler)
    clocks = sampler.stop()
    dev_ms = rank_max(dev_ms)
    frames = B * F * K * world
    value = frames / (dev_ms / 1000.0)

    # ---------------------------------------------------------------- end to end (host buffers, public API)
    ring = torch.empty(B, min(64, K * F), 3, SIZE, SIZE).pin_memory()          # host frames: a ring a consumer would drain
    pipe.render_keypoints(kps_host[:, :Wm * F], SIZE, out=ring)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.render_keypoints(kps_host[:, Wm * F:], SIZE, out=ring)
    e1.record()
    barrier()
    e2e_ms = rank_max(e0.elapsed_time(e1))
    e2e_value = frames / (e2e_ms / 1000.0)
    checksum = float(ring.double().abs().mean())
    capi.check_overflow(dev, "bench")
    h2d = B * F * 25 * 3 * 4
    d2h = B * F * 3 * SIZE * SIZE * 4

    # ---------------------------------------------------------------- roofline, other precisions, other configs (rank 0)
    roof, kernels, modes, legs = None, {}, {}, {}
    pk, pk_src = peaks()
    if rank == 0:
        roof, kernels = roofline_pass(pipe, kps_dev, B, pk, pk_src, dev_ms / 1000.0)
    del step
    if rank == 0 and not args.no_legs:
        for mode in ("strict2", "balanced", "fast"):
            if mode == args.precision:
                continue
            p2 = RenderPipeline(**PIPE_KW, precision=mode).to(dev)
            ms, _, st2 = time_steps(p2, kps_dev, B, F, 3, 1, lambda: torch.cuda.synchronize())
            modes[mode] = {"value": B * F * 3 / (ms / 1000.0), "unit": "frames/s", "ms_per_frame_step": ms / (3 * F), "parity": PARITY_NOTE[mode],
                           "timing": "device-resident, 3 steps"}
            if mode == "fast":
                r2, k2 = roofline_pass(p2, kps_dev, B, pk, pk_src, 0.0)
                modes[mode]["roofline"] = {k: r2[k] for k in ("achieved", "peak", "frac", "share_of_step", "per_class")}
                modes[mode]["roofline_memory_kernels"] = k2
            del p2, st2
            torch.cuda.empty_cache()
        # configs[0]: `bash test_start/start.sh` - batch 1, the 100 bundled keypoint frames, host keypoints in / host frames out
        kp100 = torch.from_numpy(bundled_keypoints()).unsqueeze(0).pin_memory()
        out100 = torch.empty(1, 100, 3, SIZE, SIZE).pin_memory()
        cfg0 = {}
        for mode in (args.precision, "fast"):
            p1 = pipe if mode == args.precision else RenderPipeline(**PIPE_KW, precision=mode).to(dev)
            p1.render_keypoints(kp100[:, :10], SIZE, out=out100)
            torch.cuda.synchronize()
            e0.record()
            p1.render_keypoints(kp100, SIZE, out=out100)
            e1.record()
            torch.cuda.synchronize()
            cfg0[mode] = {"value": 100.0 / (e0.elapsed_time(e1) / 1000.0), "unit": "frames/s", "ms_per_frame": e0.elapsed_time(e1) / 100.0}
        legs["configs[0]"] = {"workload": "test_start/start.sh: the 100 bundled keypoint JSONs, 512x512, batch 1 (one sequential clip), end to end "
                                          "(host keypoints in, pinned host frames out, CUDA-graph step)", "by_precision": cfg0}
        del p1
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- training legs
    train = None
    del pipe
    torch.cuda.empty_cache()
    if not args.no_train:
        from nhvr_b200.networks import define_G, define_D
        from nhvr_b200.train import UVPretrainer, RenderTrainer, synthetic_densepose, synthetic_train_batch
        TB, TS = args.train_batch, 256
        netT = define_G(3, 73, 64, "translate", 2, 5, gpu_ids=[local])
        trainer = UVPretrainer(netT, distributed=world > 1)
        pose, dp_i, dp_uv = synthetic_densepose(TB, TS, TS, dev, seed=rank)
        for _ in range(3):
            trainer.step(pose, dp_i, dp_uv)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        KT = 20
        t0.record()
        for _ in range(KT):
            last = trainer.step(pose, dp_i, dp_uv)
        t1.record()
        barrier()
        tms = rank_max(t0.elapsed_time(t1)) / KT
        eng = [e for v in netT._engines.values() if isinstance(v, list) for e in v][0]
        train = {"workload": "configs[1]: UV generator pre-train fwd+bwd+Adam, %dx%d, batch %d/GPU, 3-channel pose map (--input_nc 3, REF pretrainTrans.sh), "
                             "synthetic DensePose targets%s" % (TS, TS, TB, ", NCCL gradient all-reduce" if world > 1 else ""),
                 "steps_per_s": 1000.0 / tms, "ms_per_step": tms, "samples_per_s": world * TB * 1000.0 / tms,
                 "conv_tflops_fwd_dgrad_wgrad": 3.0 * eng.flops / (tms * 1e-3) / 1e12, "final_loss": float(last)}
        rl = train_roofline(lambda: trainer.step(pose, dp_i, dp_uv), pk)       # a training step holds collectives: every rank runs it
        if rank == 0:
            train["roofline"] = rl
        del trainer, netT, eng
        torch.cuda.empty_cache()
        if args.train_e2e_batch > 0:
            def e2e_leg(size, batch, steps):
                torch.manual_seed(0)
                kw = dict(PIPE_KW, size=size)
                pipe2 = RenderPipeline(**kw).to(dev)
                netD = define_D(PIPE_KW["pose_nc"] + 3, 64, 3, "instance", False, 2, True, gpu_ids=[local])
                tr = RenderTrainer(pipe2, netD, distributed=world > 1)
                bt = synthetic_train_batch(batch, size, dev, seed=rank)
                z = torch.zeros(batch, PIPE_KW["pose_nc"] - 3, size, size, device=dev)
                bt["pose"], bt["pose_prev"] = torch.cat([bt["pose"], z], 1), torch.cat([bt["pose_prev"], z], 1)
                for _ in range(2):
                    tr.step(bt)
                barrier()
                t0.record()
                for _ in range(steps):
                    o = tr.step(bt)
                t1.record()
                barrier()
                ems = rank_max(t0.elapsed_time(t1)) / steps
                res = {"steps_per_s": 1000.0 / ems, "ms_per_step": ems, "samples_per_s": world * batch * 1000.0 / ems,
                       "loss_G": float(o["loss_G"]), "loss_D": float(o["loss_D"])}
                rl = train_roofline(lambda: tr.step(bt), pk)                     # every rank (collectives inside)
                if rank == 0:
                    res["roofline"] = rl
                capi.check_overflow(dev, "training leg")
                return res
            EB = args.train_e2e_batch
            train["e2e_step"] = dict(workload="configs[2]: end-to-end train step 512x512, batch %d/GPU (2 frames/sample: t-1 without grad), G-side + "
                                              "multiscale PatchGAN D (ndf 64, 2 scales), Adam%s" % (EB, ", NCCL all-reduce x2" if world > 1 else ""),
                                     **e2e_leg(SIZE, EB, 5))
            torch.cuda.empty_cache()
            if not args.no_legs:
                train["e2e_step_1024"] = dict(workload="configs[4]: end-to-end train step 1024x1024, 6 pose channels, background net trained, batch 4/GPU",
                                              **e2e_leg(1024, 4, 3))
                torch.cuda.empty_cache()

    if rank == 0 and not args.no_legs:
        # configs[4] inference: 1024 x 1024, 6 pose channels, 4 lock-step clips
        kw = dict(PIPE_KW, size=1024)
        p4 = RenderPipeline(**kw, precision=args.precision).to(dev)
        st4 = p4.step_graph(4, 1024, 1024, use_graph=True)
        k4 = synthetic_keypoints(4, 12, seed=7).to(dev)
        for t in range(4):
            ops.pose_rasterize(k4[:, t].contiguous(), 1024, 6, out=st4.pose); st4.run()
        torch.cuda.synchronize()
        e0.record()
        for t in range(4, 12):
            ops.pose_rasterize(k4[:, t].contiguous(), 1024, 6, out=st4.pose); st4.run()
        e1.record()
        torch.cuda.synchronize()
        legs["configs[4]"] = {"workload": "1024x1024 rendering, 2D + LaplaceProj pose input (6 channels), refined background, 4 lock-step clips, device-resident",
                              "precision": args.precision, "value": 32.0 / (e0.elapsed_time(e1) / 1000.0), "unit": "frames/s",
                              "ms_per_frame_step": e0.elapsed_time(e1) / 8.0}
        del p4, st4
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, n, cores = cpu_oracle_fps(15.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d frame(s) of the same path at batch 1 on the host cores, fp32 torch oracle" % n}

    if rank == 0:
        out = {"metric": "frames/sec rendered @512x512", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
               "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16 operands (split precision: hi + lo, 3 MMAs per product), f32 accumulate" if args.precision == "strict"
                        else capi.operand_dtype() + " operands, f32 accumulate",
               "data": "synthetic keypoint random walk of the bundled skeleton, random-init weights (no checkpoint offline)",
               "config": workload_config(B, F, args.precision), "parity": PARITY_NOTE[args.precision], "clocks": clocks,
               "timed_region_s": dev_ms / 1000.0,
               "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "frames_checksum": checksum, "timed_region_s": e2e_ms / 1000.0},
               "gpu_launches": int(launches), "roofline": roof, "roofline_memory_kernels": kernels, "modes": modes, "legs": legs,
               "cpu_baseline": cpu, "train": train}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def train_roofline(step_fn, pk):
    """Per-kernel-kind rates of one training step (events around every native launch, un-overlapped)."""
    from nhvr_b200 import ops
    torch.cuda.synchronize()
    ops.PROFILE = []
    step_fn()
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for kind, work, a, b in recs:
        base = kind.split(":")[0]
        if base.startswith("conv"):
            base = "dgrad" if ":dgrad" in kind else "conv_fwd"
        d = agg.setdefault(base, [0.0, 0.0, 0])
        d[0] += work; d[1] += a.elapsed_time(b) * 1e-3; d[2] += 1
    out = {}
    for k, (w, s, n) in sorted(agg.items()):
        tensor = k in ("conv_fwd", "dgrad", "wgrad")
        rate = w / s / (1e12 if tensor else 1e9)
        peak = pk["bf16_tflops"] if tensor else pk["hbm_gbs"]
        out[k] = {"bound": "tensor" if tensor else "hbm", "achieved": rate, "unit": "TFLOP/s" if tensor else "GB/s", "frac": rate / peak,
                  "ms_per_step": 1e3 * s, "launches": n}
    return out


if __name__ == "__main__":
    main()
