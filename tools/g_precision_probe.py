"""CPU emulation (dev probe, imports the oracle): which roundings of the plain-fp16 temporal-generator engine produce its frame error?
Every conv of the oracle G (ngf 48 / 2 / 10, 12 input channels) is re-run with a subset of the engine's 16-bit roundings: a = conv-input activations,
w = weights, r = the raw conv output read back by the InstanceNorm apply, s = the residual stream of the ResnetBlocks.  usage: python tools/g_precision_probe.py [size]"""
import sys, os, numpy as np, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle.networks import define_G, ResnetBlock
from nhvr_b200 import pose as posemod
torch.manual_seed(0)
torch.set_num_threads(8)
SZ = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G = define_G(12, 4, 48, "temporal", 2, 10).eval()
kps = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'keypoints_body25.npy'))
pose = torch.from_numpy(posemod.pose_maps(kps[[0]], SZ, 6))
tex = torch.tanh(F.interpolate(torch.randn(1, 3, 24, 24), size=(SZ, SZ), mode="bicubic")) * (pose[:, :3].abs().sum(1, keepdim=True) > -10)
prev = torch.tanh(F.interpolate(torch.randn(1, 3, 16, 16), size=(SZ, SZ), mode="bicubic"))
x = torch.cat([tex, pose, prev], 1)
r16 = lambda t: t.half().float()

def run(mode):
    """mode flags: a = round conv-input activations to fp16, w = round weights, r = round raw conv output, s = round the residual stream"""
    mods = list(G.model)
    h = x
    i = 0
    def conv(m, inp):
        wt = r16(m.weight) if 'w' in mode else m.weight
        inp = r16(inp) if 'a' in mode else inp
        if isinstance(m, nn.ConvTranspose2d):
            return F.conv_transpose2d(inp, wt, m.bias, stride=m.stride, padding=m.padding, output_padding=m.output_padding)
        return F.conv2d(inp, wt, m.bias, stride=m.stride, padding=m.padding)
    def inorm(raw):
        mu = raw.mean((2, 3), keepdim=True); var = raw.var((2, 3), unbiased=False, keepdim=True)
        rr = r16(raw) if 'r' in mode else raw
        return (rr - mu) * torch.rsqrt(var + 1e-5)
    while i < len(mods):
        m = mods[i]
        if isinstance(m, ResnetBlock):
            cb = m.conv_block
            t = torch.relu(inorm(conv(cb[1], cb[0](h))))
            t = inorm(conv(cb[5], cb[4](t)))
            h = h + t
            if 's' in mode: h = r16(h)
            i += 1
        elif isinstance(m, nn.ReflectionPad2d):
            h = m(h); i += 1
        elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            raw = conv(m, h)
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.InstanceNorm2d):
                h = torch.relu(inorm(raw)); i += 3
            else:
                h = raw; i += 1
        else:
            h = m(h); i += 1
    return torch.cat([torch.tanh(h[:, :-1]), torch.sigmoid(h[:, -1:])], 1)

with torch.no_grad():
    ref = run('')
    for mode in ['awrs', 'awr', 'aws', 'aw', 'ars', 'wrs', 's', 'r', 'a', 'w']:
        y = run(mode)
        e = (y - ref).abs()
        print("%-5s max %.3e  99.99%% %.3e  rms %.3e" % (mode, e.max().item(), e.flatten().kthvalue(int(e.numel() * 0.9999)).values.item(), e.pow(2).mean().sqrt().item()), flush=True)
