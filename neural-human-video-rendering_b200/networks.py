"""define_G / define_D — the reference's network factory API, backed by the sm_100a kernels.

Mirrors the (absent) reference ``models/networks.py`` that BASELINE.json names: same factory names and
argument meaning as public pix2pixHD ``define_G`` / ``define_D`` [SURVEY §8(b)], same ``state_dict``
keys (``model.<idx>.weight`` / ``model.<idx>.conv_block.<j>.weight``; D: ``scale<i>_layer<j>.0.weight``)
so a real checkpoint (``<checkpoints_dir>/<name>/<epoch>_net_<label>.pth``) loads unchanged.  Widths
and depths come from the reference's flags: --ngf_global/--n_downsample_global/--n_blocks_global
[REF test_start/start.sh:15-17], --n_downsample_bg/--n_blocks_bg [REF start.sh:20-21],
--n_blocks_translate [REF pretrainTrans.sh:13].

The modules hold fp32 parameters (the checkpoint format) and a bf16 packed copy in the conv kernel's
consumption order; ``forward`` takes / returns NCHW fp32 CUDA tensors like the reference's modules and
runs entirely through the C-ABI (nhvr_b200.ops).  No torch.nn compute op is called on this path and
there is no fallback: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import capi, ops
from .capi import NhvrError


# ----------------------------------------------------------------------------------------------
# parameter holders with the reference's module indices (never called)
# ----------------------------------------------------------------------------------------------
class _Slot(nn.Module):
    """Index placeholder for a parameter-free reference module (pad / norm / activation)."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class _ConvParams(nn.Module):
    """weight/bias with nn.Conv2d (or nn.ConvTranspose2d) shapes and names."""

    def __init__(self, cin, cout, k, stride=1, pad=0, transposed=False):
        super().__init__()
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.empty(*shape).normal_(0.0, 0.02))   # pix2pixHD weights_init
        bound = 1.0 / float(cin * k * k) ** 0.5 if not transposed else 1.0 / float(cout * k * k) ** 0.5
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))  # nn.Conv2d default bias init
        self.cin, self.cout, self.k, self.stride, self.pad, self.transposed = cin, cout, k, stride, pad, transposed


class _ResnetBlockParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv_block = nn.Sequential(_Slot("ReflectionPad2d(1)"), _ConvParams(dim, dim, 3, pad=1), _Slot("InstanceNorm2d"),
                                        _Slot("ReLU"), _Slot("ReflectionPad2d(1)"), _ConvParams(dim, dim, 3, pad=1),
                                        _Slot("InstanceNorm2d"))


# ----------------------------------------------------------------------------------------------
# per-shape execution engine for a GlobalGenerator-shaped conv chain
# ----------------------------------------------------------------------------------------------
class _Step:
    __slots__ = ("plan", "params", "post", "residual_from", "act")


class _ChainEngine:
    """Buffers + conv plans of one conv chain at a fixed (N, H, W).

    chain: list of dicts {params, kind, k, stride, pad, halo, norm(bool), act, res('save'|'add'|None)}
    The last layer has norm=False and writes fp32 NCHW with bias + ``final_act``.
    """

    def __init__(self, chain: List[dict], N: int, H: int, W: int, device, final_act: int, out_channels: int, train: bool = False,
                 split3: bool = False, no_wlo: bool = False):
        """split3: split-precision inference engine (include/nhvr.h conv flag bit 3): every activation is a hilo pair,
        every K step three MMAs - fp32-class results at 3x the tensor work (no backward)."""
        assert not (train and split3), "split precision is an inference format"
        self.split3 = split3
        capi.overflow_flag(device)          # registers the range guard (fp16 overflow of a conv output / gradient)
        self.N, self.H, self.W = N, H, W
        self.train, self.final_act, self._bwd, self.busy = train, final_act, None, False
        self.device = device
        self.chain = chain
        self.plans: List[ops.ConvPlan] = []
        env = __import__("os").environ.get("NHVR_IN_FUSED")
        want_fused = (not train) and (split3 if env is None else env not in ("0", ""))       # see self.fused below
        h, w = H, W
        for i, L in enumerate(chain):
            last = i == len(chain) - 1
            p: _ConvParams = L["params"]
            if last:
                epi, act = capi.EPI_BIAS_ACT_F32, final_act
            elif L.get("norm", True):
                epi, act = capi.EPI_RAW_STATS, capi.ACT_NONE
            else:
                epi, act = capi.EPI_BIAS_ACT_P8, L["act"]        # conv + bias + activation, no norm (D layer 0)
            plan = ops.ConvPlan(capi.CONV_TRANSPOSE if p.transposed else capi.CONV, p.cin, p.cout, p.k, p.stride, p.pad,
                                N, h, w, L["halo"], epi, act, allow_tap_pairing=not train, split3=split3, no_wlo=no_wlo,
                                centred_stats=(i == 0 and self._centres_stem(chain)),
                                align_tiles=want_fused and bool(L.get("res")) and bool(__import__("os").environ.get("NHVR_ALIGN_TILES")))
            plan.label = ("head" if last else "stem" if i == 0 else "res" if L.get("res") else "up" if p.transposed
                          else "down" if p.stride == 2 else "conv") + "%dx%d_%d-%d" % (p.k, p.k, p.cin, p.cout)
            self.plans.append(plan)
            h, w = plan.Ho, plan.Wo
        self.out = torch.empty(N, out_channels, h, w, dtype=torch.float32, device=device)
        # input buffers: one per distinct input descriptor, rotating triple for residual chains
        self._bufs: Dict[tuple, List[ops.P8Buffer]] = {}
        self._raws: Dict[tuple, ops.P8Buffer] = {}
        self.in_bufs: List[ops.P8Buffer] = []
        self.raw_bufs: List[Optional[ops.P8Buffer]] = []
        live: Optional[ops.P8Buffer] = None      # residual source that a later apply still has to read
        for i, plan in enumerate(self.plans):
            d = plan.in_desc
            key = tuple(getattr(d, f) for f, _ in capi.ActDesc._fields_)
            pool = self._bufs.setdefault(key, [])
            # in_bufs[i] is written by the apply after conv i-1 (pack for i == 0): it must not be the live
            # residual source nor the buffer conv i-1 reads
            prev = self.in_bufs[-1] if self.in_bufs else None
            tapped = bool(chain[i].get("tap"))           # this layer's input is read back as a feature after the run: never reused
            chosen = None if (train or tapped) else next((b for b in pool if b is not live and b is not prev), None)
            if chosen is None:
                chosen = ops.P8Buffer(d.copy(), device)
                if not tapped:
                    pool.append(chosen)
            self.in_bufs.append(chosen)
            if i > 0 and chain[i - 1].get("res") == "add":
                live = None                          # consumed by the apply that just wrote `chosen`
            if chain[i].get("res") == "save":
                live = chosen
            if i < len(self.plans) - 1 and chain[i].get("norm", True):
                rd = plan.raw_desc()
                rkey = (rd.N, rd.C8, rd.H, rd.W, rd.hilo)
                if train:                                   # backward re-reads every layer's raw output
                    self.raw_bufs.append(ops.P8Buffer(rd, device))
                    continue
                if rkey not in self._raws:
                    self._raws[rkey] = ops.P8Buffer(rd, device)
                self.raw_bufs.append(self._raws[rkey])
            else:
                self.raw_bufs.append(None)
        # inference: layers whose InstanceNorm is finished inside the conv kernel (no raw tensor, no nhvr_in_apply launch).
        # Measured (profiles/r02b_in_fused.md): the image-wide meeting point makes co-resident CTAs run in lock-step, which
        # exposes the epilogue; split-precision layers (3x longer MMA phase, 2x larger raw tensor) gain 2.5-4 % of the frame
        # step, plain fp16 layers lose 2-7 % - so the default is split precision only.  NHVR_IN_FUSED=1 / 0 forces it on / off
        # (NHVR_NO_IN_FUSED=1 disables it inside the library as well).
        self.fused = [False] * len(self.plans)
        if want_fused:
            for i, (L, plan) in enumerate(zip(chain[:-1], self.plans[:-1])):
                self.fused[i] = bool(L.get("norm", True)) and plan.in_fused_supported()
        # InstanceNorm statistics: one zero-fill per forward; the arrival counters of the fused layers (N uint32 each) live
        # behind them in the same buffer so that the same fill re-arms them
        sizes = [N * pl.Cout8 * 8 * 4 if chain[i].get("norm", True) else 0 for i, pl in enumerate(self.plans[:-1])]   # {sum, sum sq, shift, -}
        nsync = (sum(self.fused) * N + 1) // 2
        self.stats_all = torch.zeros(max(1, sum(sizes) + nsync), dtype=torch.float64, device=device)     # fp64 sums (include/nhvr.h)
        self.stats: List[torch.Tensor] = []
        off = 0
        for s in sizes:
            self.stats.append(self.stats_all[off:off + s])
            off += s
        sync_all = self.stats_all[off:off + nsync].view(torch.int32)
        self.sync: List[Optional[torch.Tensor]] = []
        k = 0
        for f in self.fused:
            self.sync.append(sync_all[k * N:(k + 1) * N] if f else None)
            k += 1 if f else 0
        self.flops = sum(pl.flops for pl in self.plans)
        self.weight_versions: Optional[tuple] = None

    @staticmethod
    def _centres_stem(chain: List[dict]) -> bool:
        """The first layer's InstanceNorm sums are centred (ops.stem_stat_shift) when it is a normalised, narrow-input conv."""
        L0 = chain[0]
        p0 = L0["params"]
        return bool(L0.get("norm", True)) and not p0.transposed and p0.cin <= 32 and len(chain) > 1

    def pack_weights(self) -> None:
        self._pack_batch = ops.pack_weights_all(self.plans, [L["params"].weight for L in self.chain], getattr(self, "_pack_batch", None))
        if self._centres_stem(self.chain):                       # the stem's filter summed over its taps (ops.stem_stat_shift)
            w0 = self.chain[0]["params"].weight.detach().float().sum((2, 3)).contiguous()
            if getattr(self, "_wsum", None) is None:
                self._wsum = w0
            else:
                self._wsum.copy_(w0)                             # in place: a captured graph keeps reading this buffer

    def maybe_repack(self) -> None:
        ver = (tuple(L["params"].weight._version for L in self.chain) + tuple(L["params"].weight.data_ptr() for L in self.chain)
               + tuple(getattr(L["params"], "ext_version", 0) for L in self.chain))
        if ver != self.weight_versions:
            self.pack_weights()
            self.weight_versions = ver

    def run(self, inputs: Sequence[torch.Tensor]) -> torch.Tensor:
        self.stats_all.zero_()
        if self._centres_stem(self.chain):
            # centre the stem's InstanceNorm sums on its response to the flat part of the input (stick-figure pose maps)
            ops.stem_stat_shift(self._wsum, inputs, self.stats[0])
        ops.pack_nchw(inputs, self.in_bufs[0])
        return self.run_packed(zero_stats=False)

    # ------------------------------------------------------------------ backward (training engines only)
    def _build_backward(self) -> None:
        L = len(self.plans)
        dev = self.device
        B = {"wplans": [], "dplans": [None] * L, "G": [None] * L, "dX": [None] * L, "fold": [None] * L}
        gpool: Dict[tuple, ops.P8Buffer] = {}
        xpool: Dict[tuple, ops.P8Buffer] = {}

        def key_of(d):
            return tuple(getattr(d, f) for f, _ in capi.ActDesc._fields_)
        ws_bytes = 0
        for i, (Lr, plan) in enumerate(zip(self.chain, self.plans)):
            wp = ops.WgradPlan(plan)
            B["wplans"].append(wp)
            ws_bytes = max(ws_bytes, wp.ws_bytes)
            k = key_of(wp.g_desc) + ((i,) if __import__("os").environ.get("NHVR_DEBUG_KEEP") else ())
            if k not in gpool:
                gpool[k] = ops.P8Buffer(wp.g_desc.copy(), dev)
            B["G"][i] = gpool[k]
            p: _ConvParams = Lr["params"]
            d = plan.desc
            if p.transposed:          # dgrad of a transposed conv = the stride-2 conv of the gradient
                dp = ops.ConvPlan(capi.CONV, p.cout, p.cin, p.k, 2, p.pad, self.N, plan.Ho, plan.Wo, capi.HALO_ZERO, capi.EPI_RAW_P8,
                                  in_extra_rows=wp.g_desc.pad_b - p.pad)
                fold = (0, 0, False)
            elif p.stride == 2:       # dgrad of a stride-2 conv = the transposed conv of the gradient
                if (p.k, p.pad) not in ((3, 1), (4, 2)):
                    raise NhvrError("backward of a stride-2 conv is built for k3 p1 and k4 p2 only")
                dp = ops.ConvPlan(capi.CONV_TRANSPOSE, p.cout, p.cin, p.k, 2, p.pad, self.N, plan.Ho, plan.Wo, capi.HALO_ZERO,
                                  capi.EPI_RAW_P8, in_extra_cols=wp.g_desc.pad_r - 1, out_hw=(d.H, d.W))
                fold = (0, 0, False)
            else:                     # stride-1: gradient over the padded input extent, halo folded afterwards
                dp = ops.ConvPlan(capi.CONV_DGRAD_S1, p.cin, p.cout, p.k, 1, p.pad, self.N, d.H, d.W, capi.HALO_ZERO, capi.EPI_RAW_P8)
                fold = (p.pad, p.pad, Lr["halo"] == capi.HALO_REFLECT)
            ops.conv_set_input_desc(dp, wp.g_desc)
            dp.label = "dgrad_" + plan.label
            B["dplans"][i] = dp
            B["fold"][i] = fold
            xd = ops.make_desc(self.N, dp.Cout8, dp.Ho, dp.Wo)
            k = key_of(xd)
            if k not in xpool:
                xpool[k] = ops.P8Buffer(xd, dev)
            B["dX"][i] = xpool[k]
        B["ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        maxc8 = max(pl.Cout8 for pl in self.plans)
        B["sums"] = torch.zeros(self.N * maxc8 * 17, dtype=torch.float32, device=dev)      # sums + per-plane arrival counters
        B["dy"] = {}
        self._bwd = B
        self._bwd_versions = None

    def _grad_scale(self, grad_out: torch.Tensor, extra: Dict[int, torch.Tensor]) -> float:
        """Power-of-two scale that keeps the 16-bit gradient operands in range (max |g| ~ 64), WITHOUT a host
        synchronisation per step: the maximum is reduced on the device and copied to pinned host memory asynchronously;
        the scale used now comes from the most recent copy that has completed (the first call waits once).  Gradient
        magnitudes drift slowly and fp16 leaves 2^10 of headroom above 64; an overflow would raise through the range
        guard (capi.check_overflow)."""
        # max |g| in ONE reduction per tensor (abs().max() writes and re-reads a temporary: 1.1 ms per end-to-end training step)
        inf = float("inf")
        amax = torch.stack([torch.linalg.vector_norm(grad_out, inf)] + [torch.linalg.vector_norm(g, inf) for g in extra.values()]).max().float()
        st = getattr(self, "_amax_state", None)
        if st is None:
            st = self._amax_state = {"host": torch.zeros(1, dtype=torch.float32).pin_memory(), "evt": torch.cuda.Event(), "S": None,
                                     "pending": False}
        if st["pending"] and (st["S"] is None or st["evt"].query()):
            st["evt"].synchronize()
            a = float(st["host"][0])
            if a > 0.0 and math.isfinite(a):
                st["S"] = 2.0 ** round(math.log2(64.0 / a))
            st["pending"] = False
        if not st["pending"]:
            st["host"].copy_(amax.reshape(1), non_blocking=True)
            st["evt"].record()
            st["pending"] = True
        if st["S"] is None:                                  # very first backward of this engine: one blocking read
            st["evt"].synchronize()
            a = float(st["host"][0])
            st["S"] = 2.0 ** round(math.log2(64.0 / a)) if (a > 0.0 and math.isfinite(a)) else 1.0
            st["pending"] = False
        return st["S"]

    def backward(self, grad_out: Optional[torch.Tensor], need_input_grad: bool, in_channels: int,
                 extra_grads: Optional[Dict[int, torch.Tensor]] = None, need_weight_grad: bool = True):
        """Gradients of the chain: (input_grad NCHW fp32 or None, [dW per layer], [db per layer or None]).

        extra_grads[i] (i >= 1): additional gradient w.r.t. the activation that feeds conv i (the discriminator's
        intermediate features used by the feature-matching loss), NCHW fp32."""
        assert self.train, "backward needs a training engine"
        if self._bwd is None:
            self._build_backward()
        B = self._bwd
        extra_grads = {k: v for k, v in (extra_grads or {}).items() if v is not None}
        ver = tuple((Lr["params"].weight._version, getattr(Lr["params"], "ext_version", 0)) for Lr in self.chain)
        if ver != self._bwd_versions:                       # dgrad convs read the same weights, differently packed
            B["pack_batch"] = ops.pack_weights_all(B["dplans"], [Lr["params"].weight for Lr in self.chain], B.get("pack_batch"))
            self._bwd_versions = ver
        L = len(self.plans)
        if grad_out is None:
            grad_out = torch.zeros_like(self.out)
        grad_out = grad_out.contiguous().float()
        S = self._grad_scale(grad_out, extra_grads)       # exact power of two, no per-step host sync
        inv_S = 1.0 / S
        self._last_S = S
        g_pre = ops.head_bwd(self.out, grad_out, self.final_act, S)
        dbs: List[Optional[torch.Tensor]] = [None] * L
        if need_weight_grad:
            dbs[L - 1] = ops.bias_grad(g_pre, inv_S)
            b_last = self.chain[L - 1]["params"].bias
            if getattr(b_last, "_nhvr_direct_grad", False) and b_last.grad is not None:
                b_last.grad.add_(dbs[L - 1])                 # lands in the flat bucket now: the tail all-reduce below needs it
                dbs[L - 1] = None
        mid = L // 2 if (need_weight_grad and L >= 8 and getattr(self, "mid_backward", None) is not None
                         and all(Lr.get("norm", True) for Lr in self.chain[L // 2:L - 1])) else -1
        ops.pack_nchw([g_pre], B["G"][L - 1])
        dWs: List[Optional[torch.Tensor]] = [None] * L
        dy_total: Dict[int, ops.P8Buffer] = {}
        input_grad = None
        for i in range(L - 1, -1, -1):
            Lr, plan = self.chain[i], self.plans[i]
            if need_weight_grad:                             # frozen discriminator under the generator loss: input grads only
                wgt = Lr["params"].weight
                if getattr(wgt, "_nhvr_direct_grad", False) and wgt.grad is not None:
                    # train.ParamBucket: .grad is a view of the flat gradient bucket - accumulate straight into it, autograd
                    # gets None for this parameter (no allocation, no AccumulateGrad kernel, no gather / scatter copies)
                    B["wplans"][i].run(self.in_bufs[i], B["G"][i], B["ws"], wgt.grad, inv_S, accumulate=True)
                else:
                    dW = torch.empty_like(wgt, dtype=torch.float32)
                    B["wplans"][i].run(self.in_bufs[i], B["G"][i], B["ws"], dW, inv_S)
                    dWs[i] = dW
            if i == mid:
                # every weight gradient of layers >= mid is in the bucket (their biases sit in front of an InstanceNorm: zero
                # gradient, never written): reduce that half under the backward of the first half
                self.mid_backward(2 * mid)
            if i == 0 and not need_input_grad:
                break
            dX = B["dX"][i]
            B["dplans"][i].forward(B["G"][i], dX.ptr)
            pt, pl_, refl = B["fold"][i]
            if i == 0:
                d0 = self.plans[0].desc
                input_grad = ops.fold_unpack(dX, pt, pl_, refl, self.N, B["dplans"][0].Cout8, d0.H, d0.W, in_channels, inv_S)
                break
            prev = self.chain[i - 1]
            skip = dy_total.get(i + 1) if Lr.get("res") == "save" else None
            if i in extra_grads:                             # external gradient of this activation (D features)
                assert skip is None
                rd = self.plans[i - 1].raw_desc()
                key = ("extra", rd.C8, rd.H, rd.W)
                if key not in B["dy"]:
                    B["dy"][key] = ops.P8Buffer(rd, self.device)
                skip = B["dy"][key]
                ops.pack_nchw([extra_grads[i].contiguous().float() * S], skip)
            dy_out = None
            if prev.get("res") == "add":
                rd = self.plans[i - 1].raw_desc()
                slot = (i // 2) % 2                          # two live skip-gradient buffers alternate
                key = (slot, rd.C8, rd.H, rd.W)
                if key not in B["dy"]:
                    B["dy"][key] = ops.P8Buffer(rd, self.device)
                dy_out = B["dy"][key]
                dy_total[i - 1] = dy_out
            if prev.get("norm", True):
                ops.in_bwd(dX, pt, pl_, refl, self.raw_bufs[i - 1], self.stats[i - 1], prev["act"], B["sums"], B["G"][i - 1],
                           skip=skip, dy_out=dy_out)
            else:                                            # conv + bias + activation without norm
                dbias = torch.zeros(self.plans[i - 1].Cout8 * 8, dtype=torch.float32, device=self.device)
                ops.act_bwd(dX, pt, pl_, refl, self.in_bufs[i], prev["act"], B["G"][i - 1], dbias, skip=skip)
                if need_weight_grad:
                    dbs[i - 1] = dbias[:prev["params"].cout] * inv_S
        return input_grad, dWs, dbs

    def feature(self, i: int) -> torch.Tensor:
        """Activation that feeds conv i (= output of layer i-1) as NCHW fp32 (D's intermediate features)."""
        return ops.unpack_nchw(self.in_bufs[i], self.chain[i]["params"].cin)

    def run_packed(self, zero_stats: bool = True) -> torch.Tensor:
        """Run the chain assuming in_bufs[0] already holds the packed input."""
        if zero_stats:
            self.stats_all.zero_()
        res_src: Optional[ops.P8Buffer] = None
        n = len(self.plans)
        for i, (L, plan) in enumerate(zip(self.chain, self.plans)):
            x = self.in_bufs[i]
            if L.get("res") == "save":
                res_src = x
            if i == n - 1:
                plan.forward(x, self.out.data_ptr(), bias=L["params"].bias)
                break
            if not L.get("norm", True):
                nxt = self.in_bufs[i + 1]
                plan.forward(x, nxt.ptr, bias=L["params"].bias, out_desc=nxt.desc)
                continue
            if self.fused[i]:
                plan.forward_in_fused(x, self.stats[i], L["act"], self.in_bufs[i + 1], self.sync[i],
                                      residual=res_src if L.get("res") == "add" else None)
            else:
                raw = self.raw_bufs[i]
                plan.forward(x, raw.ptr, stats=self.stats[i])
                ops.in_apply(raw, self.stats[i], L["act"], self.in_bufs[i + 1],
                             residual=res_src if L.get("res") == "add" else None)
            if L.get("res") == "add":
                res_src = None
        return self.out


def _release_engine(eng) -> None:
    eng.busy = False


class _ChainFunction(torch.autograd.Function):
    """Differentiable wrapper of a conv chain: forward and backward both run on the sm_100a kernels.
    With n_feats > 0 the intermediate activations feeding convs 1..n_feats are returned too (discriminator)."""

    @staticmethod
    def forward(ctx, eng: "_ChainEngine", n_inputs: int, n_feats: int, *tensors):
        inputs = tensors[:n_inputs]
        ctx.eng, ctx.n_inputs, ctx.n_feats = eng, n_inputs, n_feats
        # a grad-enabled forward that is never back-propagated (eval without no_grad, an exception, a dropped loss term)
        # must not pin the engine: release it when autograd frees this node
        weakref.finalize(ctx, _release_engine, eng)
        ctx.in_channels = [t.shape[1] for t in inputs]
        ctx.need_in = any(t.requires_grad for t in inputs)
        out = eng.run([t.detach().float() for t in inputs]).clone()
        if n_feats == 0:
            return out
        return tuple(eng.feature(j) for j in range(1, n_feats + 1)) + (out,)

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.eng
        extra = {j + 1: g for j, g in enumerate(grads[:-1])}
        need_w = any(ctx.needs_input_grad[3 + ctx.n_inputs:])
        gin, dWs, dbs = eng.backward(grads[-1], ctx.need_in, sum(ctx.in_channels), extra, need_weight_grad=need_w)
        eng.busy = False
        grads_in = [None] * ctx.n_inputs
        if gin is not None:
            off = 0
            for k, c in enumerate(ctx.in_channels):
                grads_in[k] = gin[:, off:off + c].contiguous()
                off += c
        grads_p = []
        for i, Lr in enumerate(eng.chain):
            grads_p.append(dWs[i] if need_w else None)
            b = Lr["params"].bias
            if not need_w:
                grads_p.append(None)
            elif getattr(b, "_nhvr_direct_grad", False) and b.grad is not None:
                if dbs[i] is not None:                       # biases in front of an affine-free InstanceNorm: exactly zero, skipped
                    b.grad.add_(dbs[i])
                grads_p.append(None)
            else:
                grads_p.append(dbs[i] if dbs[i] is not None else torch.zeros_like(b))
        hook = getattr(eng, "after_backward", None)
        if hook is not None and need_w:
            hook()                                           # e.g. fire this network's gradient all-reduce (train.ParamBucket)
        return (None, None, None) + tuple(grads_in) + tuple(grads_p)


class GlobalGeneratorB200(nn.Module):
    """pix2pixHD GlobalGenerator shape on the sm_100a kernels (see module docstring).

    final: 'tanh' | 'none' | 'tanh_sigmoid_last' (SPEC D3 / D9).
    """

    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, final="tanh"):
        super().__init__()
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.n_downsampling, self.n_blocks, self.final = n_downsampling, n_blocks, final
        model: List[nn.Module] = [_Slot("ReflectionPad2d(3)"), _ConvParams(input_nc, ngf, 7, pad=3), _Slot("InstanceNorm2d"),
                                  _Slot("ReLU")]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [_ConvParams(ngf * mult, ngf * mult * 2, 3, stride=2, pad=1), _Slot("InstanceNorm2d"), _Slot("ReLU")]
        mult = 2 ** n_downsampling
        for _ in range(n_blocks):
            model += [_ResnetBlockParams(ngf * mult)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [_ConvParams(ngf * mult, ngf * mult // 2, 3, stride=2, pad=1, transposed=True),
                      _Slot("InstanceNorm2d"), _Slot("ReLU")]
        model += [_Slot("ReflectionPad2d(3)"), _ConvParams(ngf, output_nc, 7, pad=3)]
        if final == "tanh":
            model += [_Slot("Tanh")]
        self.model = nn.Sequential(*model)
        self._engines: Dict[tuple, _ChainEngine] = {}
        # inference precision: "f16" = one 16-bit operand per value (the global operand type, see capi.DEFAULT_OPERAND);
        # "split3" = split precision (hi + lo operands, 3 MMAs per K step, fp32-class results);
        # "split2" = hi + lo activations, 16-bit weights (2 MMAs per K step: x_hi*w + x_lo*w)
        self.precision = "f16"

    def set_precision(self, precision: str) -> "GlobalGeneratorB200":
        if precision not in ("f16", "split3", "split2"):
            raise NhvrError("precision must be 'f16', 'split3' or 'split2', got %r" % (precision,))
        self.precision = precision
        return self

    # ---- chain description -------------------------------------------------------------------
    def _chain(self) -> List[dict]:
        R, Z = capi.HALO_REFLECT, capi.HALO_ZERO
        chain: List[dict] = []
        mods = list(self.model)
        convs = [m for m in mods if isinstance(m, (_ConvParams, _ResnetBlockParams))]
        for m in convs:
            if isinstance(m, _ResnetBlockParams):
                chain.append(dict(params=m.conv_block[1], halo=R, act=capi.ACT_RELU, res="save"))
                chain.append(dict(params=m.conv_block[5], halo=R, act=capi.ACT_NONE, res="add"))
            else:
                chain.append(dict(params=m, halo=R if m.k == 7 else Z, act=capi.ACT_RELU, res=None))
        return chain

    def _final_act(self) -> int:
        return {"tanh": capi.ACT_TANH, "none": capi.ACT_NONE, "tanh_sigmoid_last": capi.ACT_TANH_SIGMOID_LAST}[self.final]

    def engine(self, N: int, H: int, W: int, train: bool = False) -> _ChainEngine:
        dev = self.model[1].weight.device
        if train:
            # a training engine holds the activations its backward needs: one engine per forward call still
            # waiting for its backward (e.g. two frames of one step), recycled afterwards
            pool = self._engines.setdefault((N, H, W, dev.index, "train"), [])
            eng = next((e for e in pool if not e.busy), None)
            if eng is None:
                if len(pool) >= 8:
                    raise NhvrError("more than 8 forward passes of one module are waiting for backward")
                capi.require_device()
                eng = _ChainEngine(self._chain(), N, H, W, dev, self._final_act(), self.output_nc, train=True)
                pool.append(eng)
            eng.busy = True
            eng.after_backward = getattr(self, "_after_backward", None)
            eng.mid_backward = getattr(self, "_mid_backward", None)
            eng.maybe_repack()
            return eng
        key = (N, H, W, dev.index, train, self.precision)
        eng = self._engines.get(key)
        if eng is None:
            capi.require_device()
            if dev.type != "cuda":
                raise NhvrError("GlobalGeneratorB200 parameters must live on a CUDA device (call .cuda()); no CPU path")
            eng = _ChainEngine(self._chain(), N, H, W, dev, self._final_act(), self.output_nc, train=train,
                               split3=self.precision in ("split3", "split2"), no_wlo=self.precision == "split2")
            self._engines[key] = eng
        eng.maybe_repack()
        return eng

    def forward(self, x, *more):
        """x (and optional further tensors, concatenated along C): NCHW fp32 CUDA -> NCHW fp32 CUDA."""
        xs = (x,) + tuple(more)
        for t in xs:
            if not t.is_cuda:
                raise NhvrError("nhvr_b200 modules take CUDA tensors only (no CPU fallback)")
        cin = sum(t.shape[1] for t in xs)
        if cin != self.input_nc:
            raise NhvrError("expected %d input channels, got %d" % (self.input_nc, cin))
        N, _, H, W = x.shape
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in xs)):
            eng = self.engine(N, H, W, train=True)
            params = []
            for Lr in eng.chain:
                params += [Lr["params"].weight, Lr["params"].bias]
            return _ChainFunction.apply(eng, len(xs), 0, *xs, *params)
        eng = self.engine(N, H, W)
        # the engine's output buffer is persistent (the step graph reads it in place); callers of the module API get
        # their own tensor, like every nn.Module of the reference
        return eng.run([t.float() for t in xs]).clone()


def define_G(input_nc, output_nc, ngf, netG="global", n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1,
             n_blocks_local=3, norm="instance", gpu_ids: Sequence[int] = ()):
    """pix2pixHD ``define_G`` signature [SURVEY §8(b); named by BASELINE.json].

    netG: 'global' | 'bg' (tanh), 'temporal' (RGB tanh + mask sigmoid), 'translate' (UV generator, raw 73 ch).
    """
    if norm != "instance":
        raise NotImplementedError("only norm='instance' (the reference default) is built on sm_100a")
    final = {"global": "tanh", "bg": "tanh", "temporal": "tanh_sigmoid_last", "translate": "none"}.get(netG)
    if final is None:
        raise NotImplementedError("generator [%s] not implemented" % netG)
    net = GlobalGeneratorB200(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, final=final)
    if len(gpu_ids) > 0:
        net.cuda(gpu_ids[0])
    elif torch.cuda.is_available():
        net.cuda()
    return net


# ----------------------------------------------------------------------------------------------
# multiscale PatchGAN discriminator (training side)
# ----------------------------------------------------------------------------------------------
class MultiscaleDiscriminatorB200(nn.Module):
    """pix2pixHD MultiscaleDiscriminator / NLayerDiscriminator (kw=4, padw=2) on the sm_100a kernels
    [SURVEY §8 a8, Appendix C].  Parameter names equal upstream's: ``scale<i>_layer<j>.0.weight`` with
    getIntermFeat, ``layer<i>.<idx>.weight`` without.  forward() returns the same nested lists."""

    def __init__(self, input_nc, ndf=64, n_layers=3, use_sigmoid=False, num_D=3, getIntermFeat=False):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError("use_sigmoid=True (the reference trains LSGAN: no sigmoid)")
        self.input_nc, self.ndf, self.n_layers, self.num_D, self.getIntermFeat = input_nc, ndf, n_layers, num_D, getIntermFeat
        for i in range(num_D):
            seq = self._layers(input_nc, ndf, n_layers)
            if getIntermFeat:
                for j, grp in enumerate(seq):
                    setattr(self, "scale%d_layer%d" % (i, j), nn.Sequential(*grp))
            else:
                setattr(self, "layer%d" % i, nn.Sequential(*[m for grp in seq for m in grp]))
        self._engines: Dict[tuple, List[_ChainEngine]] = {}

    @staticmethod
    def _layers(input_nc, ndf, n_layers):
        seq = [[_ConvParams(input_nc, ndf, 4, stride=2, pad=2), _Slot("LeakyReLU(0.2)")]]
        nf = ndf
        for _ in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            seq.append([_ConvParams(nf_prev, nf, 4, stride=2, pad=2), _Slot("InstanceNorm2d"), _Slot("LeakyReLU(0.2)")])
        nf_prev, nf = nf, min(nf * 2, 512)
        seq.append([_ConvParams(nf_prev, nf, 4, stride=1, pad=2), _Slot("InstanceNorm2d"), _Slot("LeakyReLU(0.2)")])
        seq.append([_ConvParams(nf, 1, 4, stride=1, pad=2)])
        return seq

    def _convs(self, scale: int) -> List[_ConvParams]:
        if self.getIntermFeat:
            mods = [m for j in range(self.n_layers + 2) for m in getattr(self, "scale%d_layer%d" % (scale, j))]
        else:
            mods = list(getattr(self, "layer%d" % scale))
        return [m for m in mods if isinstance(m, _ConvParams)]

    def _chain(self, scale: int) -> List[dict]:
        convs = self._convs(scale)
        chain = []
        for j, p in enumerate(convs):
            chain.append(dict(params=p, halo=capi.HALO_ZERO, act=capi.ACT_LRELU02, res=None, norm=(0 < j < len(convs) - 1)))
        return chain

    def _scale_engines(self, N, H, W, dev, train: bool) -> List[_ChainEngine]:
        if train:
            pool = self._engines.setdefault((N, H, W, dev.index, "train"), [])
            engs = next((e for e in pool if not any(x.busy for x in e)), None)
            if engs is None:
                if len(pool) >= 8:
                    raise NhvrError("more than 8 forward passes of one module are waiting for backward")
                engs = self._make_engines(N, H, W, dev, True)
                pool.append(engs)
            for e in engs:
                e.busy = True
                e.after_backward = getattr(self, "_after_backward", None)
            return engs
        key = (N, H, W, dev.index)
        if key not in self._engines:
            self._engines[key] = self._make_engines(N, H, W, dev, False)
        return self._engines[key]

    def _make_engines(self, N, H, W, dev, train):
        capi.require_device()
        engs, h, w = [], H, W
        for i in range(self.num_D):
            engs.append(_ChainEngine(self._chain(self.num_D - 1 - i), N, h, w, dev, capi.ACT_NONE, 1, train=train))
            h, w = (h + 1) // 2, (w + 1) // 2
        return engs

    def forward(self, x, *more):
        """x (and optional further tensors, concatenated along C: condition + image) -> per-scale lists."""
        xs = [t.contiguous().float() for t in (x,) + tuple(more)]
        for t in xs:
            if not t.is_cuda:
                raise NhvrError("nhvr_b200 modules take CUDA tensors only (no CPU fallback)")
        if sum(t.shape[1] for t in xs) != self.input_nc:
            raise NhvrError("expected %d input channels, got %d" % (self.input_nc, sum(t.shape[1] for t in xs)))
        N, _, H, W = xs[0].shape
        train = torch.is_grad_enabled() and (any(t.requires_grad for t in xs) or any(p.requires_grad for p in self.parameters()))
        engs = self._scale_engines(N, H, W, xs[0].device, train)
        result = []
        xd = xs
        for i, eng in enumerate(engs):
            eng.maybe_repack()
            n_feats = len(eng.plans) - 1 if self.getIntermFeat else 0
            if train:
                params = []
                for Lr in eng.chain:
                    params += [Lr["params"].weight, Lr["params"].bias]
                outs = _ChainFunction.apply(eng, len(xd), n_feats, *xd, *params)
                result.append(list(outs) if n_feats else [outs])
            else:
                out = eng.run(xd).clone()
                result.append([eng.feature(j) for j in range(1, n_feats + 1)] + [out])
            if i != self.num_D - 1:
                xd = [ops.avgpool3s2(t) for t in xd]
        return result



# ----------------------------------------------------------------------------------------------
# VGG19 feature stack of the perceptual loss (training side, optional)
# ----------------------------------------------------------------------------------------------
class Vgg19B200(nn.Module):
    """The five feature taps of pix2pixHD's ``Vgg19`` (relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 of torchvision
    ``vgg19().features``) on the sm_100a conv kernels - the network behind ``VGGLoss``, which pix2pixHD (README.md:101: the
    reference "borrows heavily" from it) adds to the generator objective unless ``--no_vgg_loss`` is given.
    Parameter names equal torchvision's (``features.<idx>.weight``), so ``vgg19-*.pth`` loads unchanged; there is no ImageNet
    checkpoint offline, the parity test uses random weights.  Frozen (requires_grad False): only input gradients are computed.

    One conv chain per resolution (conv + bias + ReLU, zero padding, no norm), ``nhvr_maxpool2`` between them; a stage's first
    activation is its feature tap (the chain engines expose the tensor that feeds conv 1)."""

    STAGES = ((0, 2), (5, 7), (10, 12, 14, 16), (19, 21, 23, 25), (28,))
    WIDTHS = {0: (3, 64), 2: (64, 64), 5: (64, 128), 7: (128, 128), 10: (128, 256), 12: (256, 256), 14: (256, 256), 16: (256, 256),
              19: (256, 512), 21: (512, 512), 23: (512, 512), 25: (512, 512), 28: (512, 512)}

    def __init__(self):
        super().__init__()
        mods: List[nn.Module] = []
        for idx in range(30):
            if idx in self.WIDTHS:
                cin, cout = self.WIDTHS[idx]
                m = _ConvParams(cin, cout, 3, pad=1)
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")    # torchvision's VGG init
                nn.init.zeros_(m.bias)
                mods.append(m)
            else:
                mods.append(_Slot("MaxPool2d(2, 2)" if idx in (4, 9, 18, 27) else "ReLU"))
        self.features = nn.Sequential(*mods)
        for p in self.parameters():
            p.requires_grad_(False)
        self._engines: Dict[tuple, List[List[_ChainEngine]]] = {}

    def _chain(self, stage: int) -> List[dict]:
        # the activation that feeds a stage's SECOND conv is the stage's feature tap (relu<k>_1)
        return [dict(params=self.features[i], halo=capi.HALO_ZERO, act=capi.ACT_RELU, res=None, norm=False, tap=(j == 1))
                for j, i in enumerate(self.STAGES[stage])]

    def _make_engines(self, N, H, W, dev, train) -> List[_ChainEngine]:
        capi.require_device()
        engs, h, w = [], H, W
        for s in range(5):
            cout = self.WIDTHS[self.STAGES[s][-1]][1]
            engs.append(_ChainEngine(self._chain(s), N, h, w, dev, capi.ACT_RELU, cout, train=train))
            h, w = h // 2, w // 2
        return engs

    def _stage_engines(self, N, H, W, dev, train) -> List[_ChainEngine]:
        key = (N, H, W, dev.index, train)
        pool = self._engines.setdefault(key, [])
        if train:
            engs = next((e for e in pool if not any(x.busy for x in e)), None)
            if engs is None:
                if len(pool) >= 4:
                    raise NhvrError("more than 4 forward passes of the VGG stack are waiting for backward")
                engs = self._make_engines(N, H, W, dev, True)
                pool.append(engs)
            for e in engs:
                e.busy = True
            return engs
        if not pool:
            pool.append(self._make_engines(N, H, W, dev, False))
        return pool[0]

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        x = x.contiguous().float()
        if not x.is_cuda:
            raise NhvrError("nhvr_b200 modules take CUDA tensors only (no CPU fallback)")
        N, Cc, H, W = x.shape
        if Cc != 3 or H < 16 or W < 16:
            raise NhvrError("Vgg19B200 takes [N, 3, H >= 16, W >= 16] images, got %s" % (tuple(x.shape),))
        train = torch.is_grad_enabled() and x.requires_grad
        engs = self._stage_engines(N, H, W, x.device, train)
        feats, h = [], x
        for s, eng in enumerate(engs):
            eng.maybe_repack()
            n_feats = 1 if len(eng.plans) > 1 else 0
            if train:
                params = []
                for Lr in eng.chain:
                    params += [Lr["params"].weight, Lr["params"].bias]
                outs = _ChainFunction.apply(eng, 1, n_feats, h, *params)
                feat, out = (outs[0], outs[1]) if n_feats else (outs, outs)
            else:
                out = eng.run([h]).clone()
                feat = eng.feature(1) if n_feats else out
            feats.append(feat)
            if s < 4:
                h = ops.maxpool2(out)
        return feats


def define_D(input_nc, ndf, n_layers_D, norm="instance", use_sigmoid=False, num_D=1, getIntermFeat=False,
             gpu_ids: Sequence[int] = ()):
    """pix2pixHD ``define_D`` signature [SURVEY §8(b); named by BASELINE.json]."""
    if norm != "instance":
        raise NotImplementedError("only norm='instance' (the reference default) is built on sm_100a")
    net = MultiscaleDiscriminatorB200(input_nc, ndf, n_layers_D, use_sigmoid, num_D, getIntermFeat)
    if len(gpu_ids) > 0:
        net.cuda(gpu_ids[0])
    elif torch.cuda.is_available():
        net.cuda()
    return net
