"""GPU probe at the reference's real configuration (test_start/start.sh: 512^2, 6 pose channels, bundled keypoints):
per frame, for several precision modes and both atlases,
  same-input   : frame function on the ORACLE's previous frame (pose_t, prev_{t-1}^oracle) -> out, vs the oracle
  free-running : the path's own previous frame fed back
  oracle-self  : the fp32 oracle against ITSELF with prev_0 perturbed by 1e-6 (the recurrence's own error growth)
  oracle-fp64  : the fp32 oracle against the same model evaluated in fp64 (the fp32 reference's own rounding noise)
usage: python tools/parity_probe.py [n_frames]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from nhvr_b200 import capi, pose as posemod
from nhvr_b200.pipeline import RenderPipeline
from oracle.pipeline import RenderModel

KW = dict(pose_nc=6, tex_nc=3, size=512, atlas_size=200, ngf_global=48, n_downsample_global=2, n_blocks_global=10,
          ngf_translate=64, n_downsample_translate=2, n_blocks_translate=5, ngf_bg=48, n_downsample_bg=2, n_blocks_bg=2)


def psnr(a, b, peak=2.0):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


def smooth_atlas(C, S, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(24, C, 6, 6, generator=g)
    return torch.tanh(torch.nn.functional.interpolate(low, size=(S, S), mode="bicubic", align_corners=False)).to(dev)


def main():
    dev = torch.device("cuda", 0)
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    kps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keypoints_body25.npy"))[:T]
    poses = torch.from_numpy(posemod.pose_maps(kps, 512, 6)).to(dev)
    capi.set_operand_dtype("f16")
    for atlas_kind in ("uniform", "smooth"):
        torch.manual_seed(11)
        ref = RenderModel(**KW).to(dev).eval()
        if atlas_kind == "smooth":
            with torch.no_grad():
                ref.atlas.copy_(smooth_atlas(3, 200, dev))
        with torch.no_grad():
            bg_r = ref.refine_bg()
            prev = torch.zeros(1, 3, 512, 512, device=dev)
            prev_p = prev + 1e-6 * torch.randn_like(prev)
            refs, selfdiv = [], []
            for t in range(T):
                o = ref.render_frame(poses[t:t + 1], prev, bg_r)
                op = ref.render_frame(poses[t:t + 1], prev_p, bg_r)
                prev, prev_p = o["out"], op["out"]
                refs.append(o)
                selfdiv.append((prev - prev_p).abs().max().item())
            print("[%s] oracle-self (1e-6 perturbation of prev_0): " % atlas_kind + " ".join("%.2e" % v for v in selfdiv))
            ref64 = RenderModel(**KW).to(dev).double().eval()
            ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
            bg64 = ref64.refine_bg()
            prev = torch.zeros(1, 3, 512, 512, device=dev, dtype=torch.float64)
            d64, d64_same = [], []
            for t in range(T):
                o64 = ref64.render_frame(poses[t:t + 1].double(), prev, bg64)
                prev = o64["out"]
                d64.append((refs[t]["out"].double() - prev).abs().max().item())
                o64s = ref64.render_frame(poses[t:t + 1].double(), (refs[t - 1]["out"].double() if t else torch.zeros_like(prev)), bg64)
                d64_same.append((refs[t]["out"].double() - o64s["out"]).abs().max().item())
            print("[%s] oracle fp32 vs fp64 free-running: " % atlas_kind + " ".join("%.2e" % v for v in d64))
            print("[%s] oracle fp32 vs fp64 same-input  : " % atlas_kind + " ".join("%.2e" % v for v in d64_same))
            del ref64
        modes_all = (("uv=split3,g=f16", dict(uv_precision="split3", g_precision="f16")),
                           ("uv=split3,g=split2", dict(uv_precision="split3", g_precision="split2")),
                           ("uv=split3,g=split3", dict(uv_precision="split3", g_precision="split3")),
                           ("uv=split3,g=f16,pose+1", dict(uv_precision="split3", g_precision="f16")),
                           ("uv=f16,g=f16", dict(uv_precision="f16", g_precision="f16")))
        only = os.environ.get("NHVR_PROBE_MODES")          # comma-free filter, e.g. NHVR_PROBE_MODES=split2
        for mode, prec in modes_all:
            if only and only not in mode:
                continue
            pipe = RenderPipeline(**KW, **prec).to(dev)
            pipe.load_state_dict(ref.state_dict())
            with torch.no_grad():
                bg = pipe.refine_bg()
                prev_free = torch.zeros(1, 3, 512, 512, device=dev)
                for t in range(T):
                    o = refs[t]
                    prev_ref = refs[t - 1]["out"] if t else torch.zeros_like(prev_free)
                    # "pose+1": a per-channel constant added to the input of a reflect-padded conv followed by InstanceNorm
                    # changes nothing mathematically, but removes the large common mode of the stem's accumulators
                    pin = poses[t:t + 1] + (1.0 if mode.endswith("pose+1") else 0.0)
                    r = pipe.render_frame(pin, prev_ref, bg)
                    f = pipe.render_frame(pin, prev_free, bg)
                    prev_free = f["out"]
                    print("[%s | %s] t=%d uvp %.2e part-agree %.6f tex %.2e | same-input out %.2e (%.1f dB) | free-running out %.2e (%.1f dB)"
                          % (atlas_kind, mode, t, (r["uvp"] - o["uvp"]).abs().max().item(), (r["part"] == o["part"]).float().mean().item(),
                             (r["tex"] - o["tex"]).abs().max().item(), (r["out"] - o["out"]).abs().max().item(), psnr(r["out"], o["out"]),
                             (f["out"] - o["out"]).abs().max().item(), psnr(f["out"], o["out"])))
            capi.check_overflow(dev, mode)
            del pipe
            torch.cuda.empty_cache()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
