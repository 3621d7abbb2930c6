"""One ResnetBlock conv (C -> C, 3x3, 128^2) as conv + nhvr_in_apply vs the fused conv + InstanceNorm kernel: event timings, and
the per-CTA cycle breakdown when NHVR_CONV_TRACE=1.  usage: python tools/fused_trace.py [C] [B] [f16|split3]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200 import ops, capi

C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
split3 = (sys.argv[3] if len(sys.argv) > 3 else "f16") == "split3"
H = W = int(sys.argv[4]) if len(sys.argv) > 4 else 128
dev = torch.device("cuda", 0)
capi.set_operand_dtype("f16")
ALIGN = os.environ.get("FT_ALIGN", "0") != "0"          # one tile per SM and image (conv desc flag bit 5)
plan = ops.ConvPlan(capi.CONV, C_, C_, 3, 1, 1, B, H, W, capi.HALO_REFLECT, capi.EPI_RAW_STATS, split3=split3, align_tiles=ALIGN)
plan_two = ops.ConvPlan(capi.CONV, C_, C_, 3, 1, 1, B, H, W, capi.HALO_REFLECT, capi.EPI_RAW_STATS, split3=split3)
x = torch.randn(B, C_, H, W, device=dev)
w = torch.randn(C_, C_, 3, 3, device=dev) * (1.0 / (C_ * 9) ** 0.5)
xin = ops.P8Buffer(plan.in_desc.copy(), dev); ops.pack_nchw([x], xin)
plan.pack_weights(w); plan_two.pack_weights(w)
raw = ops.P8Buffer(plan.raw_desc(), dev)
dst = ops.P8Buffer(plan.in_desc.copy(), dev)
stats = torch.zeros(B * plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
sync = torch.zeros(B, dtype=torch.int32, device=dev)
print("plan", plan.info(), "fused_supported", plan.in_fused_supported())

def two():
    stats.zero_(); plan_two.forward(xin, raw.ptr, stats=stats); ops.in_apply(raw, stats, capi.ACT_RELU, dst, residual=xin)
def fused():
    stats.zero_(); sync.zero_(); plan.forward_in_fused(xin, stats, capi.ACT_RELU, dst, sync, residual=xin)

tracing = bool(os.environ.get("NHVR_CONV_TRACE"))
only = os.environ.get("FT_ONLY")            # "fused" / "conv+apply": run just that variant 3 times (ncu captures)
if only:
    for _ in range(3):
        {"fused": fused, "conv+apply": two}[only]()
    torch.cuda.synchronize()
    sys.exit(0)
for name, fn in (("conv+apply", two), ("fused", fused)):
    for _ in range(1 if tracing else 5):
        fn()
    torch.cuda.synchronize()
    if tracing:
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-10s C=%d B=%d %s align=%d delay=%s: %.1f us per layer" % (name, C_, B, "split3" if split3 else "f16", ALIGN, os.environ.get("NHVR_FUSED_DELAY", "default"), e0.elapsed_time(e1) / 20 * 1e3))
