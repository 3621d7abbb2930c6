"""bench.py — frames/s rendered @512^2 through the hot path (BASELINE.json `metric`), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--impl reference]

A "step" advances B independent lock-step clips by one frame at 512x512 on one GPU (UV generator ->
texture lookup -> temporal generator (previous-frame conditioned) -> composite with the refined
background): B frames per step per GPU.  This is BASELINE.json configs[3]'s per-GPU shard (clip-sharded
long-sequence rendering with background compositing); configs[0] is the same path on the bundled
keypoints and is the CPU-runnable parity case.  Weights are random-init, poses synthetic (stated in `data`).

value   : frames/s with the step's poses already resident in HBM (CUDA events around each step, L2
          flushed between steps, max over ranks).
e2e     : same metric through the public API RenderPipeline.render_clips with pinned HOST poses in and
          pinned HOST frames out, the copies inside the timed region.
roofline: the tcgen05 conv kernel (dominant): algorithmic conv FLOPs of the step / summed conv-launch
          durations (CUDA events around every conv launch on the launching stream, separate pass).
cpu_baseline / --impl reference: the fp32 torch oracle (the only runnable statement of the reference's
          path — its source is absent) on the host cores, bounded sample.
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

import torch

SIZE = 512
PIPE_KW = dict(pose_nc=3, tex_nc=3, size=SIZE, atlas_size=200, ngf_global=48, n_downsample_global=2,
               n_blocks_global=10, ngf_translate=64, n_downsample_translate=2, n_blocks_translate=5, ngf_bg=48,
               n_downsample_bg=2, n_blocks_bg=2, use_mask_texture=True)


def synthetic_poses(B, T, seed=0):
    """Smooth synthetic pose maps in [-1,1] (random low-frequency blobs drifting ~5.7 px/frame; SURVEY §8d cfg 4)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, 1, 3, 16, 16, generator=g)
    drift = torch.randn(B, T, 3, 16, 16, generator=g) * 0.05
    low = base + torch.cumsum(drift, dim=1)
    up = torch.nn.functional.interpolate(low.reshape(B * T, 3, 16, 16), size=(SIZE, SIZE), mode="bilinear", align_corners=False)
    return torch.tanh(up * 2).reshape(B, T, 3, SIZE, SIZE).contiguous()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_oracle_fps(max_seconds=20.0, threads=None):
    """fp32 oracle on the host cores: frames/s over a bounded sample of the same per-frame path, batch 1."""
    from oracle.pipeline import RenderModel
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = RenderModel(**PIPE_KW).eval()
    poses = synthetic_poses(1, 4)[0]
    with torch.no_grad():
        bg = model.refine_bg()
        prev = torch.zeros(1, 3, SIZE, SIZE)
        r = model.render_frame(poses[0:1], prev, bg)      # warm-up frame
        prev = r["out"]
        n, t0 = 0, time.perf_counter()
        while True:
            r = model.render_frame(poses[(n + 1) % 4:(n + 1) % 4 + 1], prev, bg)
            prev = r["out"]
            n += 1
            el = time.perf_counter() - t0
            if el > max_seconds or n >= 8:
                break
    return n / el, n, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the oracle port on the host cores (the reference's own source is absent)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle.pipeline import RenderModel
    torch.manual_seed(0)
    model = RenderModel(**PIPE_KW).eval()
    poses = synthetic_poses(1, 4)[0]
    steps, warm = max(1, min(args.steps, 6)), max(1, min(args.warmup, 1))
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
    sample = "%d frame step(s) of the same path at batch 1 (one clip), fp32 torch oracle, %d threads" % (steps, cores)
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec rendered @512x512", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic poses, random-init weights",
        "config": workload_config(1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(clips):
    return {"workload": "configs[3] per-GPU shard: temporal clip rendering 512x512 with background compositing "
                        "(UV generator ngf64/5 blocks -> 24-part texture lookup 200^2 atlas -> temporal generator "
                        "ngf48/2 down/10 blocks -> composite)",
            "clips_in_flight_per_gpu": clips, "frames_per_step_per_gpu": clips, "resolution": SIZE,
            "l2": "flushed between timed steps (512 MiB write)", "parallelism": "clip-sharded, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=8, help="independent clips advanced in lock-step per GPU")
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_train", action="store_true", help="skip the configs[1] training leg")
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
        # stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout when the image
        # sets NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    B, K, Wm = args.clips, args.steps, args.warmup
    torch.manual_seed(0)
    pipe = RenderPipeline(**PIPE_KW).to(dev)
    T = K + Wm
    poses_host = synthetic_poses(B, T, seed=rank).pin_memory()
    poses_dev = poses_host.to(dev)
    frames_host = torch.empty(B, T, 3, SIZE, SIZE).pin_memory()
    step = pipe.step_graph(B, SIZE, SIZE, use_graph=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    step.reset()
    for t in range(Wm):
        step.pose.copy_(poses_dev[:, t]); step.run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    for i in range(K):
        flush.fill_(i & 0xFF)                      # evict L2 (not timed)
        ev[i][0].record()
        step.pose.copy_(poses_dev[:, Wm + i])
        step.run()
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms = float(tt.item())
    frames = B * K * world
    value = frames / (dev_ms / 1000.0)

    # ---------------------------------------------------------------- end to end (host buffers, public API)
    pipe.render_clips(poses_host[:, :Wm], out=frames_host[:, :Wm])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.render_clips(poses_host[:, Wm:], out=frames_host[:, Wm:])
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = frames / (float(tt.item()) / 1000.0)
    checksum = float(frames_host[:, Wm:].double().abs().mean())
    h2d = B * PIPE_KW["pose_nc"] * SIZE * SIZE * 4
    d2h = B * 3 * SIZE * SIZE * 4

    # ---------------------------------------------------------------- per-kernel roofline pass (un-graphed, events per launch)
    roof, kernels = None, {}
    if rank == 0:
        pk, pk_src = peaks()
        eager = pipe.step_graph(B, SIZE, SIZE, use_graph=False)
        eager.reset()
        for _ in range(2):
            eager.pose.copy_(poses_dev[:, 0]); eager.run()
        torch.cuda.synchronize()
        ops.PROFILE = []
        for t in range(3):
            eager.pose.copy_(poses_dev[:, t]); eager.run()
        torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for kind, work, a, b in recs:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += work; d[1] += a.elapsed_time(b) * 1e-3; d[2] += 1
        conv = agg.get("conv")
        if conv:
            ach = conv[0] / conv[1] / 1e12
            peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            traffic, traffic_src = None, None
            tp = os.path.join(ROOT, "profiles", "r01_conv_dram_traffic.json")
            if os.path.exists(tp):                      # dram__bytes_read+write per conv launch from a committed ncu capture
                tj = json.load(open(tp))
                if tj.get("clips") == B:
                    traffic, traffic_src = tj["mean_dram_bytes_per_launch"], tj["source"]
            roof = {"kernel": "conv_shiftgemm_kernel (tcgen05)", "bound": "tensor", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "frac_of_burst_peak": ach / pk["bf16_tflops"], "traffic": traffic, "traffic_unit": "bytes of DRAM per launch (mean over the step's conv launches)",
                    "traffic_source": traffic_src, "peak_source": pk_src + ", sustained",
                    "launches_per_step": conv[2] // 3, "flops_per_step": conv[0] / 3,
                    "share_of_step": conv[1] / sum(v[1] for v in agg.values())}
        for kind in ("sampler", "composite", "in_apply", "pack"):
            if kind in agg:
                w, s, n = agg[kind]
                kernels[kind] = {"bound": "hbm", "achieved": w / s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": w / s / 1e9 / pk["hbm_gbs"], "launches_per_step": n // 3,
                                 "ms_per_step": 1e3 * s / 3}

    # ---------------------------------------------------------------- training leg: configs[1] UV-generator pre-train
    train = None
    launches_per_step = step.launches_per_step
    if not args.no_train:
        from nhvr_b200.networks import define_G
        from nhvr_b200.train import UVPretrainer, synthetic_densepose
        launches_per_step = step.launches_per_step
        del step, pipe, flush
        torch.cuda.empty_cache()
        TB, TS = args.train_batch, 256
        netT = define_G(3, 73, 64, "translate", 2, 5, gpu_ids=[local])
        trainer = UVPretrainer(netT, distributed=world > 1)
        pose, dp_i, dp_uv = synthetic_densepose(TB, TS, TS, dev, seed=rank)
        for _ in range(3):
            trainer.step(pose, dp_i, dp_uv)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        KT = max(5, min(K, 20))
        t0.record()
        for _ in range(KT):
            last = trainer.step(pose, dp_i, dp_uv)
        t1.record()
        barrier()
        tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tms = float(tt.item()) / KT
        eng = [e for v in netT._engines.values() if isinstance(v, list) for e in v][0]
        fwd_flops = eng.flops
        train = {"workload": "configs[1]: UV generator pre-train fwd+bwd+Adam, %dx%d, batch %d/GPU, 3-channel pose map (--input_nc 3, REF pretrainTrans.sh), synthetic DensePose targets%s"
                             % (TS, TS, TB, ", NCCL gradient all-reduce" if world > 1 else ""),
                 "steps_per_s": 1000.0 / tms, "ms_per_step": tms, "samples_per_s": world * TB * 1000.0 / tms,
                 "conv_tflops_fwd_dgrad_wgrad": 3.0 * fwd_flops / (tms * 1e-3) / 1e12, "final_loss": float(last)}
        # ---- configs[2]: end-to-end training step (UV gen + lookup + temporal generator + multiscale PatchGAN D)
        if args.train_e2e_batch > 0:
            from nhvr_b200.networks import define_D
            from nhvr_b200.train import RenderTrainer, synthetic_train_batch
            del trainer, netT
            torch.cuda.empty_cache()
            torch.manual_seed(0)
            pipe2 = RenderPipeline(**PIPE_KW).to(dev)
            netD = define_D(6, 64, 3, "instance", False, 2, True, gpu_ids=[local])
            tr = RenderTrainer(pipe2, netD, distributed=world > 1)
            EB = args.train_e2e_batch
            batch = synthetic_train_batch(EB, SIZE, dev, seed=rank)
            for _ in range(2):
                tr.step(batch)
            barrier()
            KE = 5
            t0.record()
            for _ in range(KE):
                o = tr.step(batch)
            t1.record()
            barrier()
            tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ems = float(tt.item()) / KE
            train["e2e_step"] = {"workload": "configs[2]: end-to-end train step 512x512, batch %d/GPU (2 frames/sample: t-1 without grad), "
                                             "G-side + multiscale PatchGAN D, Adam%s" % (EB, ", NCCL all-reduce x2" if world > 1 else ""),
                                 "steps_per_s": 1000.0 / ems, "ms_per_step": ems, "samples_per_s": world * EB * 1000.0 / ems,
                                 "loss_G": float(o["loss_G"]), "loss_D": float(o["loss_D"])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, n, cores = cpu_oracle_fps(15.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d frame(s) of the same path at batch 1 on the host cores, fp32 torch oracle" % n}

    if rank == 0:
        out = {"metric": "frames/sec rendered @512x512", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
               "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": capi.operand_dtype() + " operands, f32 accumulate", "data": "synthetic poses, random-init weights (no checkpoint offline)",
               "config": workload_config(B), "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "frames_checksum": checksum},
               "gpu_launches": int(launches_per_step * K), "roofline": roof, "roofline_memory_kernels": kernels,
               "cpu_baseline": cpu, "train": train}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
