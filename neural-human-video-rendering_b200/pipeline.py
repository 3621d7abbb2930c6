"""The rendering hot path end to end on one B200, and its clip sharding across GPUs.

Call stack (inferred from REF test_start/start.sh:6-28):

    uvp  = netTransG(pose)                   UV generator              [REF pretrainTrans.sh:13]
    tex  = texture_sample(atlas, uvp)        "--TexG part"             [REF start.sh:13-14,18]
    fgm  = netG(cat(tex, pose, prev))        temporal generator        [REF start.sh:7,15-17]
    bg'  = netBG(bg)                         once per clip             [REF start.sh:12,20-21]
    out  = m*fg + (1-m)*bg'                  composite                 [REF README.md:15,52,60]
    prev <- out   (zeros at clip start, SPEC D8)

Attribute names equal oracle/pipeline.py's RenderModel so the two exchange ``state_dict``s.
Inference shards independent clips: no collective.  Several clips can advance in
lock-step as the batch dimension (``render_clips``); the per-step launch sequence is captured once in
a CUDA graph and replayed, so the host issues one graph launch per frame step.
"""
from __future__ import annotations

import contextlib
import gc
import os

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import capi, ops
from .networks import define_G

N_PARTS = 24
UV_CHANNELS = 25 + 2 * N_PARTS
PRECISION_PRESETS = {"strict": ("split3", "split3"), "strict2": ("split3", "split2"), "balanced": ("split3", "f16"), "fast": ("f16", "f16")}


@contextlib.contextmanager
def _no_gc():
    """Stream capture must not be interrupted by the cyclic garbage collector: collecting an OLD step graph (pipeline <->
    graph reference cycles are only freed by the collector) destroys its CUDA graph, which is not permitted while another
    stream is capturing and invalidates the capture.  torch.cuda.graph() collects once on entry; this keeps the automatic
    collector off for the duration of the capture."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class RenderPipeline(nn.Module):
    def __init__(self, pose_nc: int = 3, tex_nc: int = 3, size: int = 512, atlas_size: int = 200,
                 ngf_global: int = 48, n_downsample_global: int = 2, n_blocks_global: int = 10,
                 ngf_translate: int = 64, n_downsample_translate: int = 2, n_blocks_translate: int = 5,
                 ngf_bg: int = 48, n_downsample_bg: int = 2, n_blocks_bg: int = 2, use_mask_texture: bool = True,
                 precision: str = "strict", uv_precision: Optional[str] = None, g_precision: Optional[str] = None):
        """precision: inference operand precision preset (measured parity per mode in profiles/):
          "strict"   UV generator, temporal generator and background net in split precision (3 x fp16 MMAs, fp32-class):
                     the mode that meets north_star's 2e-2 / 45 dB on the reference's real configuration - default;
          "strict2"  UV generator in split precision; temporal generator and background net with split ACTIVATIONS and 16-bit weights
                     (2 MMAs per product instead of 3; measured parity in profiles/r02c_strict2.md);
          "balanced" UV generator in split precision, the others fp16: PSNR ~67 dB, max-abs ~2e-2 at the few pixels where
                     InstanceNorm of a stick-figure pose map produces |z| ~ 25-50 (fp16 is relative precision);
          "fast"     everything fp16: PSNR >= 50 dB on a texture-like atlas, UV error ~3e-2 (1.5 texels).
        uv_precision / g_precision ("f16" | "split3" | "split2") override the preset per network."""
        super().__init__()
        if precision not in PRECISION_PRESETS:
            raise ValueError("precision must be one of %s" % sorted(PRECISION_PRESETS))
        uv_precision = uv_precision or PRECISION_PRESETS[precision][0]
        g_precision = g_precision or PRECISION_PRESETS[precision][1]
        self.precision = next((k for k, v in PRECISION_PRESETS.items() if v == (uv_precision, g_precision)), "custom")
        self.pose_nc, self.tex_nc, self.size, self.atlas_size = pose_nc, tex_nc, size, atlas_size
        self.use_mask_texture = use_mask_texture
        self.netTransG = define_G(pose_nc, UV_CHANNELS, ngf_translate, "translate", n_downsample_translate,
                                  n_blocks_translate)
        self.netG = define_G(tex_nc + pose_nc + 3, 4, ngf_global, "temporal", n_downsample_global, n_blocks_global)
        self.netBG = define_G(3, 3, ngf_bg, "bg", n_downsample_bg, n_blocks_bg)
        self.uv_precision, self.g_precision = uv_precision, g_precision
        self.netTransG.set_precision(uv_precision)
        self.netG.set_precision(g_precision)
        self.netBG.set_precision(g_precision)
        self.atlas = nn.Parameter(torch.empty(N_PARTS, tex_nc, atlas_size, atlas_size).uniform_(-1, 1))
        self.bg = nn.Parameter(torch.empty(3, size, size).uniform_(-1, 1))
        self._atlas_cl: Optional[torch.Tensor] = None
        self._atlas_ver = None
        self._graphs: Dict[tuple, "_StepGraph"] = {}

    # ------------------------------------------------------------------ pieces
    def atlas_channels_last(self) -> torch.Tensor:
        ver = (self.atlas._version, self.atlas.data_ptr(), getattr(self, "ext_version", 0))
        if self._atlas_cl is None or ver != self._atlas_ver:
            self._atlas_cl = ops.atlas_to_channels_last(self.atlas)
            self._atlas_ver = ver
        return self._atlas_cl

    @torch.no_grad()
    def refine_bg(self) -> torch.Tensor:
        return self.netBG(self.bg.detach().unsqueeze(0))[0].clone()

    @torch.no_grad()
    def render_frame(self, pose: torch.Tensor, prev: torch.Tensor, bg_refined: torch.Tensor,
                     want_indices: bool = True) -> Dict[str, torch.Tensor]:
        uvp = self.netTransG(pose)
        tex, part, texel = ops.texture_sample(uvp, self.atlas_channels_last(), self.tex_nc, self.use_mask_texture,
                                              want_indices=want_indices)
        fgm = self.netG(tex, pose, prev)
        out = ops.composite(fgm, bg_refined)
        return {"out": out, "fgm": fgm, "tex": tex, "uvp": uvp, "part": part, "texel": texel}

    def forward_train(self, pose: torch.Tensor, prev: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Differentiable frame: every stage runs forward AND backward on the sm_100a kernels (autograd only
        routes the gradients).  `prev` is the (detached) previous composited frame.  Under torch.no_grad() (frame t-1 of a
        training step) the networks run their fp16 inference engines - the precision the training engines use - not the
        split-precision rendering preset."""
        if not torch.is_grad_enabled():
            nets = (self.netTransG, self.netG, self.netBG)
            saved = [n.precision for n in nets]
            try:
                for n in nets:
                    n.precision = "f16"
                return self._forward_train(pose, prev)
            finally:
                for n, p in zip(nets, saved):
                    n.precision = p
        return self._forward_train(pose, prev)

    def _forward_train(self, pose: torch.Tensor, prev: torch.Tensor) -> Dict[str, torch.Tensor]:
        uvp = self.netTransG(pose)
        tex = ops.texture_sample_diff(uvp, self.atlas, self.use_mask_texture)
        fgm = self.netG(tex, pose, prev.detach())
        bg_refined = self.netBG(self.bg.unsqueeze(0))[0]
        out = ops.composite_diff(fgm, bg_refined)
        return {"out": out, "fgm": fgm, "tex": tex, "uvp": uvp, "bg": bg_refined}

    @torch.no_grad()
    def render_clip(self, poses: torch.Tensor, use_graph: bool = True) -> torch.Tensor:
        """poses [T, pose_nc, H, W] (CUDA fp32) -> frames [T, 3, H, W]."""
        return self.render_clips(poses.unsqueeze(0), use_graph=use_graph)[0]

    @torch.no_grad()
    def render_clips(self, poses: torch.Tensor, use_graph: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """poses [B, T, pose_nc, H, W]: B independent clips advanced in lock-step -> [B, T, 3, H, W].

        ``poses`` may live in (pinned) host memory: each step's poses are copied host->device inside
        the step; ``out`` may be a pinned host tensor the frames are copied back into.
        """
        B, T = poses.shape[:2]
        H, W = poses.shape[-2:]
        dev = self.bg.device
        step = self.step_graph(B, H, W, use_graph)
        step.reset()
        if out is None:
            out = torch.empty(B, T, 3, H, W, dtype=torch.float32, device=dev)
        if poses.is_cuda and out.is_cuda:
            step.drive(T, lambda t, dst: dst.copy_(poses[:, t], non_blocking=True),
                       lambda t: out[:, t].copy_(step.out, non_blocking=True))
            return out
        step.stream_clips(poses, out)
        return out

    @torch.no_grad()
    def render_keypoints(self, kps: torch.Tensor, size: Optional[int] = None, use_graph: bool = True,
                         out: Optional[torch.Tensor] = None, src_size: float = 1024.0) -> torch.Tensor:
        """kps [B, T, 25, 3] OpenPose BODY_25 keypoints (host or device) of B lock-step clips -> frames [B, T, 3, size, size]
        (``out`` may be a pinned host tensor).  The pose maps are rasterised on the GPU inside every step."""
        size = size or self.size
        B, T = kps.shape[:2]
        dev = self.bg.device
        step = self.step_graph(B, size, size, use_graph)
        step.reset()
        if out is None:
            out = torch.empty(B, T, 3, size, size, dtype=torch.float32, device=dev)
        step.stream_keypoints(kps, out, src_size)
        return out

    def step_graph(self, B: int, H: int, W: int, use_graph: bool = True, pipelined: Optional[bool] = None) -> "_StepGraph":
        """pipelined (default: automatically for B <= 2 under a CUDA graph): the UV generator of frame t+1 runs on a second
        stream next to the temporal generator of frame t (_PipelinedStepGraph) - with one or two clips a conv launch has
        fewer tiles than the GPU has CTA slots, and only the temporal generator depends on the previous frame."""
        if pipelined is None:
            pipelined = use_graph and B * ((H + 127) // 128) * ((W + 127) // 128) <= 32
            env = os.environ.get("NHVR_PIPELINED")             # experiments: force the two-stage frame pipeline on (1) / off (0)
            if env is not None and use_graph:
                pipelined = env not in ("0", "")
        key = (B, H, W, use_graph, pipelined, self.netTransG.precision, self.netG.precision, self.netBG.precision, self.use_mask_texture)
        g = self._graphs.get(key)
        if g is None:
            g = _PipelinedStepGraph(self, B, H, W) if pipelined else _StepGraph(self, B, H, W, use_graph)
            self._graphs[key] = g
        g.refresh()           # weights / atlas / background may have changed since the graph was captured
        return g

    def state_version(self) -> tuple:
        """Changes whenever a parameter is modified in place or re-assigned (optimizer step, load_state_dict, .to())."""
        ext = sum(getattr(m, "ext_version", 0) for m in self.modules())      # native optimiser updates (train.ParamBucket)
        return tuple((p._version + ext, p.data_ptr()) for p in self.parameters())


class _StepGraph:
    """One frame step for B lock-step clips with static buffers; optionally a captured CUDA graph."""

    def __init__(self, pipe: RenderPipeline, B: int, H: int, W: int, use_graph: bool):
        capi.require_device()
        self.pipe = pipe
        dev = pipe.bg.device
        self.pose = torch.zeros(B, pipe.pose_nc, H, W, dtype=torch.float32, device=dev)
        self.prev = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev)
        self.out = self.prev          # the composite writes the next step's previous frame in place
        # persistent copies the captured graph points at: refreshed IN PLACE when the parameters change (refresh())
        self.bg_refined = pipe.refine_bg()
        self.atlas_cl = pipe.atlas_channels_last().clone()
        self.engT = pipe.netTransG.engine(B, H, W)
        self.engG = pipe.netG.engine(B, H, W)
        self._version = pipe.state_version()
        self.tex = torch.empty(B, pipe.tex_nc, H, W, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._stage = None
        self.launches_per_step = 0
        self.use_graph = use_graph
        self._capture()

    def _capture(self) -> None:
        dev = self.pose.device
        self.graph = None
        if self.use_graph:
            # warm up on a side stream (attribute set-up, weight packing), then capture
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.reset()
            g = torch.cuda.CUDAGraph()
            n0 = capi.launch_count()
            with _no_gc(), torch.cuda.graph(g):
                self._body()
            self.launches_per_step = capi.launch_count() - n0
            self.graph = g
            self.reset()
        else:
            n0 = capi.launch_count()
            self._body()
            self.launches_per_step = capi.launch_count() - n0
            self.reset()

    def reset(self) -> None:
        self.prev.zero_()

    # ---- frame protocol shared with _PipelinedStepGraph
    lookahead = False                      # True: frame t+1's pose must be in its slot BEFORE advance(t)

    def pose_slot(self, t: int) -> torch.Tensor:
        """Buffer frame t's pose maps [B, pose_nc, H, W] must be written to."""
        return self.pose

    def advance(self, t: int) -> None:
        """Render frame t (its pose is in pose_slot(t)); afterwards self.out holds the frames."""
        self.run()

    def drive(self, T: int, write_pose, emit) -> None:
        """write_pose(t, dst) enqueues frame t's pose maps into dst; emit(t) consumes self.out (frame t)."""
        self.reset()
        write_pose(0, self.pose_slot(0))
        for t in range(T):
            if self.lookahead and t + 1 < T:
                write_pose(t + 1, self.pose_slot(t + 1))
            self.advance(t)
            emit(t)
            if not self.lookahead and t + 1 < T:
                write_pose(t + 1, self.pose_slot(t + 1))

    def refresh(self) -> None:
        """Re-pack weights, re-evaluate the background net and re-copy the atlas if any parameter changed since the
        last call (training step, load_state_dict): everything is updated in the buffers the captured graph reads."""
        ver = self.pipe.state_version()
        if ver == self._version:
            return
        moved = tuple(v[1] for v in ver) != tuple(v[1] for v in self._version)
        B, H, W = self.pose.shape[0], self.pose.shape[2], self.pose.shape[3]
        self.engT = self.pipe.netTransG.engine(B, H, W)       # same cached engines; engine() re-packs changed weights
        self.engG = self.pipe.netG.engine(B, H, W)
        self.bg_refined.copy_(self.pipe.refine_bg())
        self.atlas_cl.copy_(self.pipe.atlas_channels_last())
        self._version = ver
        if moved and self.graph is not None:
            self._capture()         # parameters were re-allocated: the captured bias pointers are stale

    def _staging(self):
        if self._stage is None:
            dev = self.pose.device
            self._stage = {
                "pose": [torch.empty_like(self.pose) for _ in range(2)],
                "out": [torch.empty_like(self.out) for _ in range(2)],
                "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev),
                "pose_ready": [torch.cuda.Event() for _ in range(2)], "pose_free": [torch.cuda.Event() for _ in range(2)],
                "out_ready": [torch.cuda.Event() for _ in range(2)], "out_free": [torch.cuda.Event() for _ in range(2)],
            }
        return self._stage

    def stream_clips(self, poses: torch.Tensor, out: torch.Tensor) -> None:
        """Host-resident clips: the host->device copy of step t+1's poses and the device->host copy of step
        t-1's frames run on their own streams under step t's kernels (double-buffered device staging; each
        clip's frame is one contiguous copy, so pinned host tensors go at full PCIe rate)."""
        st = self._staging()
        B, T = poses.shape[:2]
        cur = torch.cuda.current_stream()
        h2d, d2h = st["h2d"], st["d2h"]
        h2d.wait_stream(cur)
        d2h.wait_stream(cur)
        used, pused = [False, False], [False, False]

        def fetch(t: int) -> None:
            s = t & 1
            with torch.cuda.stream(h2d):
                if pused[s]:
                    h2d.wait_event(st["pose_free"][s])
                if poses.is_cuda:
                    st["pose"][s].copy_(poses[:, t], non_blocking=True)
                else:
                    for b in range(B):
                        st["pose"][s][b].copy_(poses[b, t], non_blocking=True)
                st["pose_ready"][s].record(h2d)

        fetched = set()

        def write_pose(t: int, dst: torch.Tensor) -> None:
            s = t & 1
            if t not in fetched:
                fetch(t); fetched.add(t)
            cur.wait_event(st["pose_ready"][s])
            dst.copy_(st["pose"][s], non_blocking=True)
            st["pose_free"][s].record(cur)
            pused[s] = True
            if t + 1 < T and (t + 1) not in fetched:           # prefetch the next frame's poses under this frame's kernels
                fetch(t + 1); fetched.add(t + 1)

        def emit(t: int) -> None:
            self._emit_frame(t, out, st, used)
            used[t & 1] = True

        self.drive(T, write_pose, emit)
        cur.wait_stream(d2h)
        cur.wait_stream(h2d)
        if not out.is_cuda:
            d2h.synchronize()

    def _emit_frame(self, t: int, out: torch.Tensor, st: dict, used: list) -> None:
        """Copy step t's frames out (device staging slot t & 1, then the d2h stream) under the next step's kernels."""
        s = t & 1
        B = self.out.shape[0]
        t_out = t % out.shape[1]              # `out` shorter than the clip acts as a ring (a consumer drains it as it fills)
        cur, d2h = torch.cuda.current_stream(), st["d2h"]
        if used[s]:
            cur.wait_event(st["out_free"][s])
        st["out"][s].copy_(self.out, non_blocking=True)
        st["out_ready"][s].record(cur)
        with torch.cuda.stream(d2h):
            d2h.wait_event(st["out_ready"][s])
            if out.is_cuda:
                out[:, t_out].copy_(st["out"][s], non_blocking=True)
            else:
                for b in range(B):
                    out[b, t_out].copy_(st["out"][s][b], non_blocking=True)
            st["out_free"][s].record(d2h)

    def stream_keypoints(self, kps: torch.Tensor, out: torch.Tensor, src_size: float = 1024.0) -> None:
        """Clips given as OpenPose keypoints [B, T, 25, 3] (device): every step rasterises its pose maps on the GPU
        (ops.pose_rasterize, bit-identical to nhvr_b200.pose) straight into the step's input - 300 bytes per frame cross
        PCIe instead of a 3-6 MB pose map - and streams the frames out like stream_clips."""
        st = self._staging()
        B, T = kps.shape[:2]
        H = self.pose.shape[-1]
        cur = torch.cuda.current_stream()
        st["d2h"].wait_stream(cur)
        used = [False, False]
        kps = kps.to(self.pose.device, torch.float32)
        nc = self.pose.shape[1]

        def emit(t: int) -> None:
            self._emit_frame(t, out, st, used)
            used[t & 1] = True

        self.drive(T, lambda t, dst: ops.pose_rasterize(kps[:, t].contiguous(), H, nc, src_size, out=dst), emit)
        cur.wait_stream(st["d2h"])
        if not out.is_cuda:
            st["d2h"].synchronize()

    def _body(self) -> None:
        pipe = self.pipe
        uvp = self.engT.run([self.pose])
        ops.texture_sample(uvp, self.atlas_cl, pipe.tex_nc, pipe.use_mask_texture, tex_out=self.tex, want_indices=False)
        fgm = self.engG.run([self.tex, self.pose, self.prev])
        ops.composite(fgm, self.bg_refined, out=self.prev)

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()


class _PipelinedStepGraph(_StepGraph):
    """Frame step for ONE or TWO clips: a two-stage software pipeline across frames inside one CUDA graph.

    Only the temporal generator depends on the previous frame; the UV generator and the texture lookup of frame t+1 need
    nothing but its pose.  With a single clip every conv launch has fewer tiles (130 at 128^2) than the GPU has CTA slots
    (296), so the graph of step t forks into two streams: netG(tex_t, pose_t, prev) -> composite, and next to it
    netTransG(pose_{t+1}) -> lookup -> tex_{t+1}; the two nets' CTAs share the SMs.  Two graphs alternate (buffer parity)."""

    lookahead = True

    def __init__(self, pipe: "RenderPipeline", B: int, H: int, W: int):
        capi.require_device()
        self.pipe = pipe
        dev = pipe.bg.device
        self.poses = [torch.zeros(B, pipe.pose_nc, H, W, dtype=torch.float32, device=dev) for _ in range(2)]
        self.texs = [torch.zeros(B, pipe.tex_nc, H, W, dtype=torch.float32, device=dev) for _ in range(2)]
        self.pose = self.poses[0]              # shape / device carrier for the staging helpers
        self.prev = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev)
        self.out = self.prev
        self.bg_refined = pipe.refine_bg()
        self.atlas_cl = pipe.atlas_channels_last().clone()
        self.engT = pipe.netTransG.engine(B, H, W)
        self.engG = pipe.netG.engine(B, H, W)
        self._version = pipe.state_version()
        self._stage = None
        self.use_graph = True
        self.side = torch.cuda.Stream(device=dev)
        self.graph = None
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.launches_per_step = 0
        self._capture()

    def _uv_stage(self, p: int) -> None:
        """UV generator + lookup for the pose in slot p -> texs[p]."""
        pipe = self.pipe
        uvp = self.engT.run([self.poses[p]])
        ops.texture_sample(uvp, self.atlas_cl, pipe.tex_nc, pipe.use_mask_texture, tex_out=self.texs[p], want_indices=False)

    def _g_stage(self, p: int) -> None:
        fgm = self.engG.run([self.texs[p], self.poses[p], self.prev])
        ops.composite(fgm, self.bg_refined, out=self.prev)

    def _frame(self, p: int) -> None:
        """Frame in slot p is rendered while the frame in slot 1-p gets its texture (fork / join on the side stream)."""
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            self._uv_stage(1 - p)
        self._g_stage(p)
        cur.wait_stream(self.side)

    def _capture(self) -> None:
        dev = self.prev.device
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                              # warm-up: attribute set-up, weight packing
            self._uv_stage(0)
            for p in (0, 1):
                self._frame(p)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graphs = []
        for p in (0, 1):
            g = torch.cuda.CUDAGraph()
            n0 = capi.launch_count()
            with _no_gc(), torch.cuda.graph(g):
                self._frame(p)
            self.launches_per_step = capi.launch_count() - n0
            self.graphs.append(g)
        self.graph = self.graphs[0]
        self.reset()

    def pose_slot(self, t: int) -> torch.Tensor:
        return self.poses[t & 1]

    def advance(self, t: int) -> None:
        if t == 0:
            self._uv_stage(0)                                   # pipeline prologue: frame 0's texture
        self.graphs[t & 1].replay()

    def run(self) -> None:
        raise capi.NhvrError("the pipelined step is driven through advance(t) / drive()")


# ------------------------------------------------------------------ clip sharding (no collective)
def shard_frames(n_frames: int, world_size: int, rank: int, clips_per_rank: int = 1) -> List[Tuple[int, int]]:
    """Contiguous clip ranges [(start, stop), ...] of a length-n_frames sequence owned by ``rank``.

    The sequence is cut into world_size*clips_per_rank contiguous clips (the first ``rem`` one frame
    longer); rank r owns clips r*clips_per_rank .. (r+1)*clips_per_rank-1.  Each clip restarts the
    previous-frame state at zeros (SPEC D8), so parity is defined per clip.
    """
    n_clips = world_size * clips_per_rank
    base, rem = divmod(n_frames, n_clips)
    bounds = [0]
    for c in range(n_clips):
        bounds.append(bounds[-1] + base + (1 if c < rem else 0))
    return [(bounds[c], bounds[c + 1]) for c in range(rank * clips_per_rank, (rank + 1) * clips_per_rank)]
