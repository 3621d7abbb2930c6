"""Generates the committed golden fixtures under tests/golden/ (run from the repo root, CPU only):

  ref_flags.json      the flag lists of the reference's five launch scripts, parsed verbatim from
                      /root/reference (the only pinned interface)
  keypoints_frame0.json  one OpenPose fixture copied from /root/reference/keypoints (input format pin)
  oracle_small.npz    seeded inputs + outputs of the fp32 oracle (networks / texture / composite / losses)

The reference holds no golden vectors or tests of its own — parity is UNPINNED; these
vectors pin the ORACLE against regressions and against its independent numpy restatement.
"""
import json
import os
import re
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def script_flags(path):
    txt = open(path).read()
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    txt = txt.replace("\\\n", " ")
    cmds = [l for l in txt.splitlines() if re.search(r"python3?\s+\S+\.py", l)]
    out = []
    for c in cmds:
        toks = c.split()
        i = next(k for k, t in enumerate(toks) if t.endswith(".py"))
        out.append({"entry": toks[i].lstrip("./"), "argv": toks[i + 1:]})   # a trailing lone "\\" (pretrainTrans.sh:16, EOF) is kept verbatim
    return out


def main():
    flags = {}
    for rel in ["test_start/start.sh", "train_start/pretrain_start.sh", "pretrainTrans.sh", "pre_train_tex.sh",
                "data/data_prep/run_alignPose.sh"]:
        flags[rel] = script_flags(os.path.join(REF, rel))
    json.dump(flags, open(os.path.join(OUT, "ref_flags.json"), "w"), indent=1)
    shutil.copy(os.path.join(REF, "keypoints", "frame00000_keypoints.json"), os.path.join(OUT, "keypoints_frame0.json"))

    from oracle.networks import define_G, define_D
    from oracle.texture import texture_sample, composite
    from oracle import losses
    torch.manual_seed(1234)
    g = define_G(5, 4, 8, "temporal", 1, 1).eval()
    x = torch.randn(1, 5, 16, 16)
    d = define_D(7, 8, 2, "instance", False, 2, True).eval()
    xd = torch.randn(1, 7, 24, 24)
    uvp = torch.randn(1, 73, 5, 6)
    atlas = torch.rand(24, 3, 8, 8)
    fgm = torch.rand(2, 4, 5, 6)
    bg = torch.rand(3, 5, 6)
    with torch.no_grad():
        y = g(x)
        yd = d(xd)
        tex, part, texel = texture_sample(uvp, atlas, True)
        comp = composite(fgm, bg)
        dp_i = torch.randint(0, 25, (1, 5, 6))
        dp_uv = torch.rand(1, 2, 5, 6)
        l_uv = losses.uv_loss(uvp, dp_i, dp_uv)
        l_prob = losses.prob_loss(uvp, dp_i)
        l_gan = losses.gan_loss(yd, True)
    sd = {("g." + k): v.numpy() for k, v in g.state_dict().items()}
    sd.update({("d." + k): v.numpy() for k, v in d.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "oracle_small.npz"), x=x.numpy(), y=y.numpy(), xd=xd.numpy(),
                        yd_last0=yd[0][-1].numpy(), yd_last1=yd[1][-1].numpy(), yd_feat00=yd[0][0].numpy(),
                        uvp=uvp.numpy(), atlas=atlas.numpy(), tex=tex.numpy(), part=part.numpy(), texel=texel.numpy(),
                        fgm=fgm.numpy(), bg=bg.numpy(), comp=comp.numpy(), dp_i=dp_i.numpy(), dp_uv=dp_uv.numpy(),
                        l_uv=l_uv.numpy(), l_prob=l_prob.numpy(), l_gan=l_gan.numpy(), **sd)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
