"""Thin Python operators over the C-ABI (include/nhvr.h): P8 buffers, conv plans, InstanceNorm apply,
texture lookup, composite.  torch is used only to own device memory and to name the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import capi
from .capi import ActDesc, ConvDesc, check, load, ptr, stream_ptr


# bench.py's per-kernel roofline pass: when PROFILE is a list, every launch below is bracketed by CUDA
# events on the launching stream and recorded as (kind, algorithmic work, start, stop).
PROFILE = None


class _prof:
    __slots__ = ("kind", "work", "e0")

    def __init__(self, kind: str, work: float):
        self.kind, self.work, self.e0 = kind, work, None

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, self.work, self.e0, e1))
        return False


def _act_interior_bytes(d: ActDesc) -> float:
    return float(d.N) * d.C8 * d.H * d.W * 16.0          # hilo: C8 is physical, both halves are really moved


class P8Buffer:
    """A P8 (planar-by-8 bf16, halo-padded) activation in device memory."""

    def __init__(self, desc: ActDesc, device=None, zero: bool = True):
        self.desc = desc
        nbytes = load().nhvr_act_bytes(C.byref(desc))
        alloc = torch.zeros if zero else torch.empty
        self.mem = alloc(nbytes, dtype=torch.uint8, device=device or torch.device("cuda"))

    @property
    def ptr(self) -> int:
        return self.mem.data_ptr()


def make_desc(N, C8, H, W, pad=(0, 0, 0, 0), split=0, halo=capi.HALO_ZERO, hilo=0) -> ActDesc:
    """hilo=1: split-precision activation (include/nhvr.h): pass the LOGICAL plane count, C8 becomes physical."""
    d = ActDesc()
    d.N, d.C8, d.H, d.W = N, C8 * (2 if hilo else 1), H, W
    d.pad_t, d.pad_l, d.pad_b, d.pad_r = pad
    d.split, d.halo, d.hilo = split, halo, hilo
    return d


def pack_nchw(srcs: Sequence[torch.Tensor], dst: P8Buffer) -> None:
    """cat(srcs, dim=1) fp32 NCHW -> P8 bf16 with halo (nhvr_pack_nchw)."""
    n = len(srcs)
    arr = (C.c_void_p * n)()
    cs = (C.c_int32 * n)()
    keep = []
    for i, s in enumerate(srcs):
        s = s.contiguous()
        assert s.dtype == torch.float32 and s.is_cuda and s.dim() == 4
        assert s.shape[0] == dst.desc.N and s.shape[2] == dst.desc.H and s.shape[3] == dst.desc.W, \
            (tuple(s.shape), dst.desc.N, dst.desc.H, dst.desc.W)
        keep.append(s)
        arr[i] = s.data_ptr()
        cs[i] = s.shape[1]
    work = sum(float(k.numel()) * 4.0 for k in keep) + _act_interior_bytes(dst.desc)
    with _prof("pack", work):
        check(load().nhvr_pack_nchw(arr, cs, n, dst.ptr, C.byref(dst.desc), stream_ptr()), "nhvr_pack_nchw")


def stem_stat_shift(wsum: torch.Tensor, srcs: Sequence[torch.Tensor], stats: torch.Tensor) -> None:
    """nhvr_stem_stat_shift: centre the (zeroed) statistics record of a first layer on the conv output of the flat input.
    wsum [Cout, Cin] = the stem's filter summed over its taps."""
    n = len(srcs)
    arr = (C.c_void_p * n)()
    cs = (C.c_int32 * n)()
    keep = [t.contiguous() for t in srcs]
    for i, t in enumerate(keep):
        assert t.dtype == torch.float32 and t.is_cuda
        arr[i] = t.data_ptr()
        cs[i] = t.shape[1]
    assert wsum.is_contiguous() and wsum.dtype == torch.float32 and wsum.dim() == 2
    N, _, H, W = keep[0].shape
    check(load().nhvr_stem_stat_shift(wsum.data_ptr(), wsum.shape[0], wsum.shape[1], arr, cs, n, N, H, W,
                                      stats.data_ptr(), stream_ptr()), "nhvr_stem_stat_shift")


def unpack_nchw(src: P8Buffer, channels: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    d = src.desc
    if out is None:
        out = torch.empty(d.N, channels, d.H, d.W, dtype=torch.float32, device=src.mem.device)
    check(load().nhvr_unpack_nchw(src.ptr, C.byref(d), out.data_ptr(), channels, stream_ptr()), "nhvr_unpack_nchw")
    return out


class ConvPlan:
    """One conv layer lowered to the tcgen05 shift-GEMM kernel (nhvr_conv_plan_*)."""

    def __init__(self, kind, Cin, Cout, k, stride, pad, N, H, W, halo, epilogue, act=capi.ACT_NONE, in_extra_rows=0,
                 in_extra_cols=0, out_hw=(0, 0), allow_tap_pairing: bool = False, split3: bool = False, centred_stats: bool = False, no_wlo: bool = False,
                 align_tiles: bool = False):
        """allow_tap_pairing: flag bit 2 of nhvr_conv_desc - the input may use the single-plane tap-paired format
        (only when no wgrad plan reads the same input buffer, i.e. inference engines).
        split3: flag bit 3 - split precision (hilo input / weights, three MMAs per K step, hilo RAW output).
        no_wlo: flag bit 6 (with split3) - no w_lo blocks: two MMAs per K step, weights rounded to 16 bits."""
        d = ConvDesc()
        d.flags = (4 if allow_tap_pairing and not split3 else 0) | (8 if split3 else 0) | (16 if centred_stats else 0) | (32 if align_tiles else 0) | (64 if (no_wlo and split3) else 0)
        self.split3 = split3
        d.in_extra_rows, d.in_extra_cols = in_extra_rows, in_extra_cols
        d.out_h, d.out_w = out_hw
        d.kind, d.Cin, d.Cout, d.kh, d.kw, d.stride, d.pad = kind, Cin, Cout, k, k, stride, pad
        d.N, d.H, d.W, d.halo, d.epilogue, d.act = N, H, W, halo, epilogue, act
        self.desc = d
        h = C.c_void_p()
        check(load().nhvr_conv_plan_create(C.byref(d), C.byref(h)), "nhvr_conv_plan_create")
        self.handle = h
        self.in_desc = ActDesc()
        check(load().nhvr_conv_input_desc(h, C.byref(self.in_desc)), "nhvr_conv_input_desc")
        ho, wo, c8 = C.c_int32(), C.c_int32(), C.c_int32()
        check(load().nhvr_conv_output_dims(h, C.byref(ho), C.byref(wo), C.byref(c8)), "nhvr_conv_output_dims")
        self.Ho, self.Wo, self.Cout8 = ho.value, wo.value, c8.value
        self.Cout, self.Cin, self.N = Cout, Cin, N
        self.epilogue = epilogue
        self.label = "conv"                 # layer class for bench.py's per-class roofline (set by the engines)
        self.weight_bytes = load().nhvr_conv_weight_bytes(h)
        self.flops = load().nhvr_conv_flops(h)
        self.packed: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().nhvr_conv_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def info(self) -> dict:
        arr = (C.c_int32 * 16)()
        check(load().nhvr_conv_plan_info(self.handle, arr, 16))
        keys = ("kcp", "nchunks", "njobs", "nruns", "nacc", "slab_units", "Npad", "bpb", "nbstages", "SA", "SB",
                "tmem_cols", "smem_bytes", "tiles_per_img", "nsplit", "nblocks")
        return dict(zip(keys, list(arr)))

    def raw_desc(self) -> ActDesc:
        """Descriptor of the RAW_STATS output (un-padded P8; a hilo activation for split-precision plans)."""
        return make_desc(self.N, self.Cout8, self.Ho, self.Wo, hilo=1 if self.split3 else 0)

    def pack_weights(self, w: torch.Tensor) -> torch.Tensor:
        w = w.detach().contiguous().float()
        assert w.is_cuda
        if self.packed is None:
            self.packed = torch.empty(self.weight_bytes, dtype=torch.uint8, device=w.device)
        check(load().nhvr_conv_pack_weights(self.handle, w.data_ptr(), self.packed.data_ptr(), stream_ptr()),
              "nhvr_conv_pack_weights")
        return self.packed

    def forward(self, x: P8Buffer, out_ptr: int, bias: Optional[torch.Tensor] = None,
                out_desc: Optional[ActDesc] = None, stats: Optional[torch.Tensor] = None) -> None:
        assert self.packed is not None, "pack_weights() first"
        with _prof(("conv3:" if self.split3 else "conv:") + self.label, self.flops):
            check(load().nhvr_conv_forward(self.handle, x.ptr, self.packed.data_ptr(), ptr(bias), out_ptr,
                                           C.byref(out_desc) if out_desc is not None else None, ptr(stats), stream_ptr()),
                  "nhvr_conv_forward")


    def in_fused_supported(self) -> bool:
        """conv + InstanceNorm + activation (+ residual) + halo in one kernel is possible for this plan on this GPU
        (include/nhvr.h nhvr_conv_in_fused_supported: every CTA of an image resident at once)."""
        return bool(load().nhvr_conv_in_fused_supported(self.handle))

    def forward_in_fused(self, x: P8Buffer, stats: torch.Tensor, act: int, dst: P8Buffer, sync: torch.Tensor,
                         residual: Optional[P8Buffer] = None, eps: float = 1e-5) -> None:
        """RAW_STATS plan: the InstanceNorm apply happens in the conv's epilogue, `dst` (the consumer's input buffer) is
        written directly.  sync: int32 [N] zeros (this launch's arrival counters, re-zeroed by the caller before the next use)."""
        assert self.packed is not None, "pack_weights() first"
        with _prof(("conv3:" if self.split3 else "conv:") + self.label, self.flops):
            check(load().nhvr_conv_forward_in_fused(self.handle, x.ptr, self.packed.data_ptr(), stats.data_ptr(), eps, act,
                                                    residual.ptr if residual is not None else None,
                                                    C.byref(residual.desc) if residual is not None else None,
                                                    dst.ptr, C.byref(dst.desc), sync.data_ptr(), stream_ptr()),
                  "nhvr_conv_forward_in_fused")


class PackBatch:
    """All conv weights of a network packed by ONE launch (nhvr_conv_pack_weights_batched).  Built for a fixed set of
    (plan, fp32 weight) pairs: the per-layer records hold raw device pointers, so the batch is valid as long as the weights'
    data_ptr() do not change (train.ParamBucket keeps parameters in one flat buffer that Adam updates in place)."""

    def __init__(self, plans: Sequence["ConvPlan"], weights: Sequence[torch.Tensor]):
        lib = load()
        rb = lib.nhvr_conv_pack_record_bytes()
        host = (C.c_uint8 * (rb * len(plans)))()
        self.ptrs = tuple(w.data_ptr() for w in weights)
        self.max_units = 0
        for i, (plan, w) in enumerate(zip(plans, weights)):
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
            if plan.packed is None:
                plan.packed = torch.empty(plan.weight_bytes, dtype=torch.uint8, device=w.device)
            units = C.c_int64()
            check(lib.nhvr_conv_pack_record_fill(plan.handle, w.data_ptr(), plan.packed.data_ptr(), C.byref(host, i * rb), C.byref(units)),
                  "nhvr_conv_pack_record_fill")
            self.max_units = max(self.max_units, units.value)
        self.n = len(plans)
        self.records = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(weights[0].device)

    def valid_for(self, weights: Sequence[torch.Tensor]) -> bool:
        return self.ptrs == tuple(w.data_ptr() for w in weights)

    def run(self) -> None:
        check(load().nhvr_conv_pack_weights_batched(self.records.data_ptr(), self.n, self.max_units, stream_ptr()),
              "nhvr_conv_pack_weights_batched")


def pack_weights_all(plans: Sequence["ConvPlan"], weights: Sequence[torch.Tensor], cache: Optional["PackBatch"]) -> Optional["PackBatch"]:
    """Pack every layer: one launch when the plans allow it (plain 16-bit operands, contiguous fp32 CUDA weights), else layer by layer.
    Returns the batch to cache (None when the per-layer path was taken)."""
    ws = [w.detach() for w in weights]
    batchable = all((not pl.split3) for pl in plans) and all(w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() for w in ws)
    if not batchable or __import__("os").environ.get("NHVR_PACK_BATCH") == "0":
        for pl, w in zip(plans, ws):
            pl.pack_weights(w)
        return None
    if cache is None or not cache.valid_for(ws) or cache.n != len(plans):
        cache = PackBatch(plans, ws)
    cache.run()
    return cache


class WgradPlan:
    """Weight-gradient plan of one forward conv (nhvr_wgrad_*): dW from the forward's P8 input and the output
    gradient stored in ``g_desc`` (which the matching dgrad conv also reads)."""

    def __init__(self, fwd: "ConvPlan"):
        h = C.c_void_p()
        check(load().nhvr_wgrad_plan_create(C.byref(fwd.desc), C.byref(h)), "nhvr_wgrad_plan_create")
        self.handle = h
        self.g_desc = ActDesc()
        check(load().nhvr_wgrad_grad_desc(h, C.byref(self.g_desc)), "nhvr_wgrad_grad_desc")
        self.ws_bytes = load().nhvr_wgrad_workspace_bytes(h)
        self.flops = fwd.flops

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().nhvr_wgrad_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def run(self, x: P8Buffer, g: P8Buffer, ws: torch.Tensor, dw: torch.Tensor, scale: float, accumulate: bool = False) -> None:
        assert dw.dtype == torch.float32 and dw.is_contiguous() and ws.numel() >= self.ws_bytes
        with _prof("wgrad", self.flops):
            check(load().nhvr_wgrad(self.handle, x.ptr, g.ptr, ws.data_ptr(), dw.data_ptr(), scale, int(accumulate), stream_ptr()),
                  "nhvr_wgrad")


def conv_set_input_desc(plan: "ConvPlan", desc: ActDesc) -> None:
    check(load().nhvr_conv_plan_set_input_desc(plan.handle, C.byref(desc)), "nhvr_conv_plan_set_input_desc")
    plan.in_desc = desc.copy()


def in_bwd(dx: P8Buffer, pad_t: int, pad_l: int, reflect: bool, raw: P8Buffer, stats: torch.Tensor, act: int, sums: torch.Tensor,
           g: P8Buffer, skip: Optional[P8Buffer] = None, dy_out: Optional[P8Buffer] = None, eps: float = 1e-5) -> None:
    work = _act_interior_bytes(raw.desc) * (4.0 + (1.0 if skip is not None else 0.0))
    with _prof("in_bwd", work):
        check(load().nhvr_in_bwd(dx.ptr, dx.desc.H, dx.desc.W, pad_t, pad_l, int(reflect), skip.ptr if skip is not None else None,
                                 raw.ptr, C.byref(raw.desc), stats.data_ptr(), eps, act, sums.data_ptr(), g.ptr, C.byref(g.desc),
                                 dy_out.ptr if dy_out is not None else None, stream_ptr()), "nhvr_in_bwd")


def act_bwd(dx: P8Buffer, pad_t: int, pad_l: int, reflect: bool, yact: P8Buffer, act: int, g: P8Buffer, dbias: torch.Tensor,
            skip: Optional[P8Buffer] = None) -> None:
    check(load().nhvr_act_bwd(dx.ptr, dx.desc.H, dx.desc.W, pad_t, pad_l, int(reflect), skip.ptr if skip is not None else None,
                              yact.ptr, C.byref(yact.desc), act, g.ptr, C.byref(g.desc), dbias.data_ptr(), stream_ptr()), "nhvr_act_bwd")


class _AvgPoolFn(torch.autograd.Function):
    """AvgPool2d(3, 2, 1, count_include_pad=False) between discriminator scales, forward and backward native."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        N, Cc, H, W = x.shape
        out = torch.empty(N, Cc, (H + 1) // 2, (W + 1) // 2, dtype=torch.float32, device=x.device)
        check(load().nhvr_avgpool3s2(x.data_ptr(), N, Cc, H, W, out.data_ptr(), stream_ptr()), "nhvr_avgpool3s2")
        ctx.shape = (N, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        N, Cc, H, W = ctx.shape
        g = g.contiguous().float()
        gin = torch.empty(N, Cc, H, W, dtype=torch.float32, device=g.device)
        check(load().nhvr_avgpool3s2_bwd(g.data_ptr(), N, Cc, H, W, 0, gin.data_ptr(), stream_ptr()), "nhvr_avgpool3s2_bwd")
        return gin


def avgpool3s2(x: torch.Tensor) -> torch.Tensor:
    return _AvgPoolFn.apply(x)


class _MaxPool2Fn(torch.autograd.Function):
    """MaxPool2d(2, 2) between the VGG19 stages (perceptual loss), forward and backward native."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        N, Cc, H, W = x.shape
        out = torch.empty(N, Cc, H // 2, W // 2, dtype=torch.float32, device=x.device)
        check(load().nhvr_maxpool2(x.data_ptr(), N, Cc, H, W, out.data_ptr(), stream_ptr()), "nhvr_maxpool2")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        N, Cc, H, W = x.shape
        g = g.contiguous().float()
        gin = torch.empty_like(x)
        check(load().nhvr_maxpool2_bwd(x.data_ptr(), g.data_ptr(), N, Cc, H, W, gin.data_ptr(), stream_ptr()), "nhvr_maxpool2_bwd")
        return gin


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    return _MaxPool2Fn.apply(x)


def fold_unpack(dx: P8Buffer, pad_t: int, pad_l: int, reflect: bool, N: int, C8: int, H: int, W: int, channels: int,
                scale: float) -> torch.Tensor:
    out = torch.empty(N, channels, H, W, dtype=torch.float32, device=dx.mem.device)
    d = make_desc(N, C8, H, W)
    check(load().nhvr_fold_unpack(dx.ptr, dx.desc.H, dx.desc.W, pad_t, pad_l, int(reflect), C.byref(d), out.data_ptr(), channels,
                                  scale, stream_ptr()), "nhvr_fold_unpack")
    return out


def head_bwd(out: torch.Tensor, grad_out: torch.Tensor, act: int, scale: float) -> torch.Tensor:
    N, Cc, H, W = out.shape
    g = torch.empty_like(out)
    check(load().nhvr_head_bwd(out.data_ptr(), grad_out.data_ptr(), N, Cc, H, W, act, scale, g.data_ptr(), stream_ptr()),
          "nhvr_head_bwd")
    return g


def bias_grad(g: torch.Tensor, scale: float) -> torch.Tensor:
    N, Cc, H, W = g.shape
    db = torch.empty(Cc, dtype=torch.float32, device=g.device)
    check(load().nhvr_bias_grad(g.data_ptr(), N, Cc, H, W, scale, 0, db.data_ptr(), stream_ptr()), "nhvr_bias_grad")
    return db


def in_apply(raw: P8Buffer, stats: torch.Tensor, act: int, dst: P8Buffer, residual: Optional[P8Buffer] = None,
             eps: float = 1e-5) -> None:
    # algorithmic bytes: read raw (+ residual), write the interior once
    work = _act_interior_bytes(raw.desc) * (3.0 if residual is not None else 2.0)
    with _prof("in_apply", work):
        check(load().nhvr_in_apply(raw.ptr, C.byref(raw.desc), stats.data_ptr(), eps, act,
                                   residual.ptr if residual is not None else None,
                                   C.byref(residual.desc) if residual is not None else None,
                                   dst.ptr, C.byref(dst.desc), stream_ptr()), "nhvr_in_apply")


def atlas_to_channels_last(atlas: torch.Tensor) -> torch.Tensor:
    """[24, Ctex, S, S] fp32 -> [24, S, S, Ct4] fp32 (the gather-friendly layout nhvr_texture_sample reads)."""
    P, Ct, S, _ = atlas.shape
    ct4 = (Ct + 3) // 4 * 4
    out = torch.zeros(P, S, S, ct4, dtype=torch.float32, device=atlas.device)
    out[..., :Ct] = atlas.detach().float().permute(0, 2, 3, 1)
    return out.contiguous()


def texture_sample(uvp: torch.Tensor, atlas_cl: torch.Tensor, Ctex: int, use_mask_texture: bool = True,
                   tex_out: Optional[torch.Tensor] = None, want_indices: bool = True):
    """nhvr_texture_sample: returns (tex [N,Ctex,H,W] f32, part [N,H,W] u8, texel [N,H,W,2] i16).
    atlas_cl: channels-last [24, S, S, Ct4] (ops.atlas_to_channels_last)."""
    assert uvp.is_cuda and uvp.dtype == torch.float32 and uvp.is_contiguous() and uvp.shape[1] == 73
    N, _, H, W = uvp.shape
    S = atlas_cl.shape[1]
    if tex_out is None:
        tex_out = torch.empty(N, Ctex, H, W, dtype=torch.float32, device=uvp.device)
    part = torch.empty(N, H, W, dtype=torch.uint8, device=uvp.device) if want_indices else None
    texel = torch.empty(N, H, W, 2, dtype=torch.int16, device=uvp.device) if want_indices else None
    # algorithmic bytes: 73 fp32 channels read + Ctex fp32 written per pixel; atlas L2-resident, excluded
    with _prof("sampler", float(N) * H * W * (73 + Ctex) * 4.0):
        check(load().nhvr_texture_sample(uvp.data_ptr(), atlas_cl.data_ptr(), N, H, W, S, Ctex, int(use_mask_texture),
                                         tex_out.data_ptr(), ptr(part), ptr(texel), stream_ptr()), "nhvr_texture_sample")
    return tex_out, part, texel


def composite(fgm: torch.Tensor, bg: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nhvr_composite: out = m*fg + (1-m)*bg."""
    assert fgm.is_cuda and fgm.dtype == torch.float32 and fgm.is_contiguous() and fgm.shape[1] == 4
    assert bg.is_cuda and bg.dtype == torch.float32 and bg.is_contiguous()
    N, _, H, W = fgm.shape
    batched = int(bg.dim() == 4 and bg.shape[0] == N and N > 1)
    if out is None:
        out = torch.empty(N, 3, H, W, dtype=torch.float32, device=fgm.device)
    # algorithmic bytes: (3 fg + 1 mask + 3 bg) read + 3 written, fp32, bg counted per frame
    with _prof("composite", float(N) * H * W * 10 * 4.0):
        check(load().nhvr_composite(fgm.data_ptr(), bg.data_ptr(), batched, N, H, W, out.data_ptr(), stream_ptr()),
              "nhvr_composite")
    return out


# ----------------------------------------------------------------------------------------------
# differentiable wrappers (training): forward and backward both on the sm_100a kernels
# ----------------------------------------------------------------------------------------------
class _TextureSampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uvp, atlas, use_mask_texture):
        u = uvp.detach().contiguous().float()
        acl = atlas_to_channels_last(atlas)
        tex, _, _ = texture_sample(u, acl, atlas.shape[1], bool(use_mask_texture), want_indices=False)
        ctx.saved = (u, acl, atlas.shape, bool(use_mask_texture))
        return tex

    @staticmethod
    def backward(ctx, gtex):
        u, acl, ashape, use_mask = ctx.saved
        N, _, H, W = u.shape
        P, Ct, S, _ = ashape
        gtex = gtex.contiguous().float()
        guvp = torch.empty_like(u)
        gacl = torch.zeros_like(acl)
        with _prof("sampler_bwd", float(N) * H * W * (73 * 2 + Ct) * 4.0):
            check(load().nhvr_texture_sample_bwd(u.data_ptr(), acl.data_ptr(), gtex.data_ptr(), N, H, W, S, Ct, int(use_mask),
                                                 guvp.data_ptr(), gacl.data_ptr(), stream_ptr()), "nhvr_texture_sample_bwd")
        gatlas = gacl[..., :Ct].permute(0, 3, 1, 2).contiguous()
        return guvp, gatlas, None


def texture_sample_diff(uvp: torch.Tensor, atlas: torch.Tensor, use_mask_texture: bool = True) -> torch.Tensor:
    """Differentiable texture lookup: tex [N,Ctex,H,W]; gradients flow to uvp and to the atlas parameter."""
    return _TextureSampleFn.apply(uvp, atlas, use_mask_texture)


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fgm, bg):
        f = fgm.detach().contiguous().float()
        b = bg.detach().contiguous().float()
        ctx.saved = (f, b)
        return composite(f, b)

    @staticmethod
    def backward(ctx, gout):
        f, b = ctx.saved
        N, _, H, W = f.shape
        batched = int(b.dim() == 4 and b.shape[0] == N and N > 1)
        gout = gout.contiguous().float()
        gf = torch.empty_like(f)
        gb = torch.empty_like(b)
        check(load().nhvr_composite_bwd(f.data_ptr(), b.data_ptr(), batched, gout.data_ptr(), N, H, W, gf.data_ptr(), gb.data_ptr(),
                                        stream_ptr()), "nhvr_composite_bwd")
        return gf, gb


def composite_diff(fgm: torch.Tensor, bg: torch.Tensor) -> torch.Tensor:
    return _CompositeFn.apply(fgm, bg)


def pose_rasterize(kps: torch.Tensor, size: int, pose_nc: int = 3, src_size: float = 1024.0, thickness: float = 4.0,
                   conf_thresh: float = 0.05, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nhvr_pose_rasterize: device keypoints [T,25,3] fp32 -> pose maps [T,pose_nc,size,size] fp32 (stick figure in [-1,1],
    bit-identical to nhvr_b200.pose.rasterize; Laplace channels zero)."""
    from .pose import LIMB_COLORS
    assert kps.is_cuda and kps.dtype == torch.float32 and kps.shape[1:] == (25, 3)
    kps = kps.contiguous()
    T = kps.shape[0]
    if out is None:
        out = torch.empty(T, pose_nc, size, size, dtype=torch.float32, device=kps.device)
    cols = (C.c_uint8 * 72)(*[int(v) for v in LIMB_COLORS.reshape(-1)])
    check(load().nhvr_pose_rasterize(kps.data_ptr(), T, size, float(src_size), float(thickness), float(conf_thresh), pose_nc,
                                     C.cast(cols, C.c_void_p), out.data_ptr(), stream_ptr()), "nhvr_pose_rasterize")
    return out


class TextureUnfolder:
    """unfold_texture on the GPU (README.md:64): accumulate frames + DensePose IUV into the 24 x S x S part atlas
    (nhvr_texture_unfold, the lookup's adjoint), then divide by the accumulated weights."""

    def __init__(self, S: int = 200, C_: int = 3, device=None):
        self.S, self.C = S, C_
        self.Q = (C_ + 1 + 3) // 4 * 4
        self.acc = torch.zeros(24, S, S, self.Q, dtype=torch.float32, device=device or torch.device("cuda"))

    def add(self, img: torch.Tensor, dp_i: torch.Tensor, dp_uv: torch.Tensor) -> None:
        img, dp_uv = img.contiguous().float(), dp_uv.contiguous().float()
        dp_i = dp_i.contiguous().to(torch.int32)
        N, Cc, H, W = img.shape
        assert Cc == self.C and dp_i.shape == (N, H, W) and dp_uv.shape == (N, 2, H, W)
        check(load().nhvr_texture_unfold(img.data_ptr(), dp_i.data_ptr(), dp_uv.data_ptr(), N, H, W, self.S, self.C, self.acc.data_ptr(),
                                         stream_ptr()), "nhvr_texture_unfold")

    def atlas(self, min_weight: float = 0.0) -> torch.Tensor:
        out = torch.empty(24, self.C, self.S, self.S, dtype=torch.float32, device=self.acc.device)
        check(load().nhvr_texture_unfold_finish(self.acc.data_ptr(), self.S, self.C, float(min_weight), out.data_ptr(), stream_ptr()),
              "nhvr_texture_unfold_finish")
        return out
