"""Texture lookup kernel on three kinds of UV input (NHVR_SAMPLER_MINB = 8: 64-register variant; default 80 registers).
usage: python tools/sampler_bench.py [B] [size]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SZ = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)
atlas = torch.rand(24, 3, 200, 200, device=dev) * 2 - 1
acl = ops.atlas_to_channels_last(atlas)
cases = {}
cases["white-noise UV"] = torch.randn(B, 73, SZ, SZ, device=dev)
low = torch.randn(B, 73, SZ // 32, SZ // 32, device=dev)
cases["smooth UV"] = torch.nn.functional.interpolate(low, size=(SZ, SZ), mode="bilinear", align_corners=False).contiguous()
flat = torch.zeros(B, 73, SZ, SZ, device=dev); flat[:, :, 200:300, 240:270] = cases["white-noise UV"][:, :, 200:300, 240:270]
cases["stick-figure-like (98 % flat)"] = flat
algo = B * SZ * SZ * (73 + 3) * 4.0
for name, uvp in cases.items():
    res = {}
    for label, a in (("plain atlas", acl),):
        for _ in range(3):
            tex, part, texel = ops.texture_sample(uvp, a, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.texture_sample(uvp, a, 3, tex_out=tex, want_indices=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[label] = (tex.clone(), part.clone(), texel.clone())
        print("%-32s %-12s %8.1f us  %7.1f GB/s algorithmic  (MINB=%s)" % (name, label, ms * 1e3, algo / ms / 1e6, os.environ.get("NHVR_SAMPLER_MINB", "6")))

# backward (nhvr_texture_sample_bwd through the C-ABI): NHVR_SAMPLER_BWD_AGG=0 disables the warp-aggregated atlas reductions
from nhvr_b200.capi import load, check
from nhvr_b200.ops import stream_ptr
gtex = torch.randn(B, 3, SZ, SZ, device=dev)
for name, uvp in cases.items():
    guvp = torch.empty_like(uvp); gacl = torch.zeros_like(acl)
    run = lambda: check(load().nhvr_texture_sample_bwd(uvp.data_ptr(), acl.data_ptr(), gtex.data_ptr(), B, SZ, SZ, 200, 3, 1,
                                                       guvp.data_ptr(), gacl.data_ptr(), stream_ptr()), "nhvr_texture_sample_bwd")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    print("%-32s backward     %8.1f us  (AGG=%s)  atlas grad checksum %.6e" % (name, e0.elapsed_time(e1) / 5 * 1e3, os.environ.get("NHVR_SAMPLER_BWD_AGG", "1"),
                                                                              gacl.double().sum().item() / 7))
