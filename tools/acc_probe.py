"""Is the tcgen05 fp32 accumulation round-to-nearest?  One conv whose operands are exactly representable in fp16 and whose
products are all positive (the accumulator grows monotonically): signed relative error of the fp32 result against fp64,
as a function of the number of accumulation steps (K = Cin * 9 / 16 MMAs per output)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nhvr_b200 import capi, ops

dev = torch.device("cuda", 0)
capi.set_operand_dtype("f16")
for cin in (16, 64, 256, 1024):
    for signed in (False, True):
        g = torch.Generator().manual_seed(cin)
        x = (torch.rand(1, cin, 40, 44, generator=g) + 0.5).half().float()
        w = ((torch.rand(16, cin, 3, 3, generator=g) + 0.5) / 64).half().float()
        if signed:
            x = x * torch.sign(torch.randn(x.shape, generator=g))
        x, w = x.to(dev), w.to(dev)
        plan = ops.ConvPlan(capi.CONV, cin, 16, 3, 1, 1, 1, 40, 44, capi.HALO_REFLECT, capi.EPI_BIAS_ACT_F32, capi.ACT_NONE)
        xin = ops.P8Buffer(plan.in_desc.copy(), dev)
        ops.pack_nchw([x], xin)
        plan.pack_weights(w)
        out = torch.empty(1, 16, 40, 44, device=dev)
        plan.forward(xin, out.data_ptr())
        ref64 = torch.nn.functional.conv2d(torch.nn.functional.pad(x.double(), (1,) * 4, mode="reflect"), w.double())
        ref32 = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1,) * 4, mode="reflect"), w)
        scale = ref64.abs().mean()
        e = (out.double() - ref64) / scale
        e32 = (ref32.double() - ref64) / scale
        print("Cin %4d (%3d MMAs/output) %s: tcgen05 mean signed err %+.3e  rms %.3e | cuDNN fp32 mean %+.3e rms %.3e   (1 ulp = 6e-8)"
              % (cin, cin * 9 // 16, "signed  " if signed else "positive", e.mean().item(), e.pow(2).mean().sqrt().item(),
                 e32.mean().item(), e32.pow(2).mean().sqrt().item()))
