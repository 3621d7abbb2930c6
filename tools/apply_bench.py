"""Micro-benchmark of the split-precision IN-apply pass (NHVR_APPLY_HILO_VARIANT = 1 / 2 / 3): GB/s of moved bytes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200 import capi, ops
dev = torch.device("cuda", 0)
capi.set_operand_dtype("f16")
for (N, C8, H, res, hilo) in ((8, 24, 128, True, 1), (8, 32, 128, True, 1), (8, 32, 128, False, 1), (8, 8, 512, False, 1), (8, 24, 128, True, 0), (8, 6, 512, False, 0)):
    raw = ops.P8Buffer(ops.make_desc(N, C8, H, H, hilo=hilo), dev)
    dst = ops.P8Buffer(ops.make_desc(N, C8, H, H, (1, 1, 1, 1), 0, capi.HALO_REFLECT, hilo=hilo), dev)
    rs = ops.P8Buffer(ops.make_desc(N, C8, H, H, (1, 1, 1, 1), 0, capi.HALO_REFLECT, hilo=hilo), dev) if res else None
    raw.mem.random_(0, 60)
    stats = torch.zeros(N * C8 * 8 * 4, dtype=torch.float64, device=dev)
    stats.view(-1, 4)[:, 1] = H * H
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for v in ((1, 2, 3) if hilo else (0,)):
        os.environ["NHVR_APPLY_HILO_VARIANT"] = str(v)
        ts = []
        for i in range(8):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.in_apply(raw, stats, capi.ACT_RELU, dst, residual=rs)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        byts = N * C8 * (2 if hilo else 1) * H * H * 16.0 * (3 if res else 2)
        print("N%d C%d %d^2 res=%d hilo=%d variant %d: %.1f us  %.0f GB/s" % (N, C8 * 8, H, res, hilo, v, ms * 1e3, byts / ms / 1e6))
