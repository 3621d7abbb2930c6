"""Time one conv layer and print the kernel's per-CTA cycle breakdown (NHVR_CONV_TRACE).
usage: python tools/conv_trace.py kind cin cout k stride pad N H W halo(R|Z) epi(RAW_STATS|BIAS_ACT_F32) [split3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nhvr_b200 import capi, ops

a = sys.argv[1:]
kind = getattr(capi, a[0])
cin, cout, k, stride, pad, N, H, W = map(int, a[1:9])
halo = capi.HALO_REFLECT if a[9] == "R" else capi.HALO_ZERO
epi = getattr(capi, "EPI_" + a[10])
split3 = len(a) > 11 and a[11] == "split3"
dev = torch.device("cuda", 0)
plan = ops.ConvPlan(kind, cin, cout, k, stride, pad, N, H, W, halo, epi, capi.ACT_TANH_SIGMOID_LAST if cout == 4 else capi.ACT_NONE,
                    split3=split3, allow_tap_pairing=True)
print(plan.info())
x = torch.rand(N, cin, H, W, device=dev) * 2 - 1
transposed = kind == capi.CONV_TRANSPOSE
w = torch.randn(*((cin, cout, k, k) if transposed else (cout, cin, k, k)), device=dev) * 0.02
xin = ops.P8Buffer(plan.in_desc.copy(), dev)
ops.pack_nchw([x], xin)
plan.pack_weights(w)
if epi == capi.EPI_RAW_STATS:
    out = ops.P8Buffer(plan.raw_desc(), dev)
    stats = torch.zeros(N * plan.Cout8 * 8 * 4, dtype=torch.float64, device=dev)
    run = lambda: plan.forward(xin, out.ptr, stats=stats)
else:
    o = torch.empty(N, cout, plan.Ho, plan.Wo, device=dev)
    run = lambda: plan.forward(xin, o.data_ptr())
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("%.1f us  %.1f TFLOP/s (algorithmic)" % (ms * 1e3, plan.flops / ms / 1e9))
os.environ["NHVR_CONV_TRACE"] = "1"
run()
torch.cuda.synchronize()
