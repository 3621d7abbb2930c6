"""Texture lookup kernels: 4 x LDG.128 per (pixel, part) on the channels-last atlas vs 2 x LDG.256 on the pair atlas
(NHVR_SAMPLER_MINB = 4 | 6 | 8 selects the pair kernel's register budget).  usage: python tools/sampler_bench.py [B] [size]"""
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
ap = ops.atlas_pair(acl)
cases = {}
cases["white-noise UV"] = torch.randn(B, 73, SZ, SZ, device=dev)
low = torch.randn(B, 73, SZ // 32, SZ // 32, device=dev)
cases["smooth UV"] = torch.nn.functional.interpolate(low, size=(SZ, SZ), mode="bilinear", align_corners=False).contiguous()
flat = torch.zeros(B, 73, SZ, SZ, device=dev); flat[:, :, 200:300, 240:270] = cases["white-noise UV"][:, :, 200:300, 240:270]
cases["stick-figure-like (98 % flat)"] = flat
algo = B * SZ * SZ * (73 + 3) * 4.0
for name, uvp in cases.items():
    res = {}
    for label, a in (("plain atlas", acl), ("pair atlas", ap)):
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
        print("%-32s %-12s %8.1f us  %7.1f GB/s algorithmic  (MINB=%s)" % (name, label, ms * 1e3, algo / ms / 1e6, os.environ.get("NHVR_SAMPLER_MINB", "4")))
    a, b = res["plain atlas"], res["pair atlas"]
    print("   plain vs pair: tex max diff %.2e, part equal %s, texel equal %s" % ((a[0] - b[0]).abs().max().item(), torch.equal(a[1], b[1]), torch.equal(a[2], b[2])))
