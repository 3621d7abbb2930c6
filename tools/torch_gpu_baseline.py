"""Stock PyTorch / cuDNN on the same B200 for the same frame step - the "no custom kernel" GPU baseline that
SURVEY.md section 8(d) recommends timing next to the CPU oracle.  Plain torch.nn modules written here (this file does
not import oracle/ or the nhvr package): UV generator -> 24-part texture lookup (grid_sample) -> temporal generator ->
composite, B clips in lock-step at 512x512, random weights.

usage: python tools/torch_gpu_baseline.py [clips] [steps]     (prints one JSON line per precision mode)
"""
import json
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F


def generator(cin, cout, ngf, n_down, n_blocks):
    L = [nn.ReflectionPad2d(3), nn.Conv2d(cin, ngf, 7), nn.InstanceNorm2d(ngf), nn.ReLU(True)]
    for i in range(n_down):
        m = 2 ** i
        L += [nn.Conv2d(ngf * m, ngf * m * 2, 3, stride=2, padding=1), nn.InstanceNorm2d(ngf * m * 2), nn.ReLU(True)]

    class Block(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.b = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(d, d, 3), nn.InstanceNorm2d(d), nn.ReLU(True),
                                   nn.ReflectionPad2d(1), nn.Conv2d(d, d, 3), nn.InstanceNorm2d(d))

        def forward(self, x):
            return x + self.b(x)
    d = ngf * 2 ** n_down
    L += [Block(d) for _ in range(n_blocks)]
    for i in range(n_down):
        m = 2 ** (n_down - i)
        L += [nn.ConvTranspose2d(ngf * m, ngf * m // 2, 3, stride=2, padding=1, output_padding=1), nn.InstanceNorm2d(ngf * m // 2),
              nn.ReLU(True)]
    L += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, cout, 7)]
    return nn.Sequential(*L)


def texture_lookup(uvp, atlas):
    """soft 24-part blend with bilinear taps (grid_sample, align_corners=True == u*(S-1) texel coordinates)."""
    p = torch.softmax(uvp[:, :25].float(), 1)
    u = (uvp[:, 25:49].float() * 0.5 + 0.5).clamp(0, 1) * 2 - 1
    v = (uvp[:, 49:73].float() * 0.5 + 0.5).clamp(0, 1) * 2 - 1
    out = 0
    for k in range(24):
        grid = torch.stack([u[:, k], v[:, k]], -1)
        t = F.grid_sample(atlas[k:k + 1].expand(uvp.shape[0], -1, -1, -1), grid, mode="bilinear", padding_mode="border", align_corners=True)
        out = out + p[:, k + 1:k + 2] * t
    return out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    netT = generator(3, 73, 64, 2, 5).to(dev).eval()
    netG = generator(9, 4, 48, 2, 10).to(dev).eval()
    atlas = (torch.rand(24, 3, 200, 200, device=dev) * 2 - 1)
    bg = torch.rand(1, 3, 512, 512, device=dev) * 2 - 1
    pose = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
    torch.backends.cudnn.benchmark = True

    def step(prev, dtype, cl):
        with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
            x = pose.contiguous(memory_format=torch.channels_last) if cl else pose
            uvp = netT(x)
            tex = texture_lookup(uvp, atlas)
            inp = torch.cat([tex.to(pose.dtype), pose, prev], 1)
            if cl:
                inp = inp.contiguous(memory_format=torch.channels_last)
            y = netG(inp).float()
        rgb, m = torch.tanh(y[:, :3]), torch.sigmoid(y[:, 3:4])
        return m * rgb + (1 - m) * bg

    for name, dtype, cl, tf32 in [("fp32 (TF32 off)", None, False, False), ("fp32 + TF32", None, False, True),
                                  ("bf16 autocast, channels_last", torch.bfloat16, True, True),
                                  ("fp16 autocast, channels_last", torch.float16, True, True)]:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        if cl:
            netT.to(memory_format=torch.channels_last); netG.to(memory_format=torch.channels_last)
        with torch.no_grad():
            prev = torch.zeros(B, 3, 512, 512, device=dev)
            for _ in range(3):
                prev = step(prev, dtype, cl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                prev = step(prev, dtype, cl)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"baseline": "stock torch %s / cuDNN, eager" % torch.__version__, "mode": name, "clips": B, "ms_per_step": ms,
                          "frames_per_s": B / ms * 1e3}))


if __name__ == "__main__":
    main()
