"""Where does the split-precision UV generator lose accuracy?  Per-layer comparison of the engine's activations (the
input of conv i, unpacked from its hilo P8 buffer) with the fp32 oracle's and with an fp64 evaluation of the oracle,
on a stick-figure pose map of the bundled keypoints at 512^2.  usage: python tools/layer_probe.py [f16|split3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from nhvr_b200 import capi, ops, pose as posemod
from nhvr_b200.networks import define_G
from oracle.networks import define_G as oracle_define_G, ResnetBlock

dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "split3"
capi.set_operand_dtype("f16")
torch.manual_seed(11)
ref = oracle_define_G(6, 73, 64, "translate", 2, 5).to(dev).eval()
ref64 = oracle_define_G(6, 73, 64, "translate", 2, 5).to(dev).double().eval()
ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
net = define_G(6, 73, 64, "translate", 2, 5)
net.load_state_dict(ref.state_dict())
net.set_precision(prec)
kps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keypoints_body25.npy"))[:1]
x = torch.from_numpy(posemod.pose_maps(kps, 512, 6)).to(dev)


def conv_inputs(model, inp):
    """inputs of every Conv2d / ConvTranspose2d in execution order (before its ReflectionPad2d)"""
    acts = []
    mods = list(model.model)
    h = inp
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, ResnetBlock):
            acts.append(h)                      # input of the block's first conv
            cb = list(m.conv_block)
            t = cb[3](cb[2](cb[1](cb[0](h))))
            acts.append(t)                      # input of the second conv
            h = h + cb[6](cb[5](cb[4](t)))
        elif isinstance(m, torch.nn.ReflectionPad2d):
            acts.append(h)
            h = m(h)
        elif isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            if not isinstance(mods[i - 1], torch.nn.ReflectionPad2d):
                acts.append(h)
            h = m(h)
        else:
            h = m(h)
        i += 1
    return acts, h


with torch.no_grad():
    a32, y32 = conv_inputs(ref, x)
    a64, y64 = conv_inputs(ref64, x.double())
    y = net(x)
    eng = net.engine(1, 512, 512)
    print("layer : |act| max / rms | engine vs fp64 max, rms | oracle-fp32 vs fp64 max, rms")
    for i in range(1, len(eng.plans)):
        cin = eng.chain[i]["params"].cin
        e = ops.unpack_nchw(eng.in_bufs[i], cin).double()
        r64 = a64[i]
        d, d32 = (e - r64), (a32[i].double() - r64)
        print("%2d  %4d ch %4d^2 : %8.2f %6.3f | %.2e %.2e | %.2e %.2e" % (i, cin, e.shape[-1], r64.abs().max().item(), r64.pow(2).mean().sqrt().item(),
              d.abs().max().item(), d.pow(2).mean().sqrt().item(), d32.abs().max().item(), d32.pow(2).mean().sqrt().item()))
    d, d32 = (y.double() - y64), (y32.double() - y64)
    print("out 73 ch : |y| max %.2f rms %.3f | engine vs fp64 max %.2e rms %.2e | oracle-fp32 vs fp64 max %.2e rms %.2e"
          % (y64.abs().max().item(), y64.pow(2).mean().sqrt().item(), d.abs().max().item(), d.pow(2).mean().sqrt().item(),
             d32.abs().max().item(), d32.pow(2).mean().sqrt().item()))
    idx = d.abs().flatten().argmax().item()
    c, yy, xx = idx // (512 * 512), (idx // 512) % 512, idx % 512
    print("worst output element: channel %d at (y %d, x %d), value %.3f" % (c, yy, xx, y64.flatten()[idx].item()))
    q = torch.quantile(d.abs().flatten()[::37].float(), torch.tensor([0.5, 0.99, 0.9999], device=dev))
    print("abs err quantiles 50%% %.2e 99%% %.2e 99.99%% %.2e" % tuple(q.tolist()))
