"""GPU probe: per-tensor gradient error of the conv-chain backward vs torch autograd on the fp32 oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from nhvr_b200.networks import define_G
from oracle.networks import define_G as oracle_define_G

dev = torch.device("cuda", 0)
args = (5, 3, 16, "global", 1, 1) if len(sys.argv) < 2 else eval(sys.argv[1])
size = 48 if len(sys.argv) < 3 else int(sys.argv[2])
batch = 2
torch.manual_seed(31)
ref = oracle_define_G(*args).to(dev)
net = define_G(*args)
net.load_state_dict(ref.state_dict())
torch.manual_seed(32)
x = (torch.rand(batch, args[0], size, size, device=dev) * 2 - 1).requires_grad_(True)
xr = x.detach().clone().requires_grad_(True)
wgt = torch.randn(batch, args[1], size, size, device=dev)
y = net(x)
yr = ref(xr)
print("fwd err", (y - yr).abs().max().item())
(y * wgt).mean().backward()
(yr * wgt).mean().backward()
rows = [("input", x.grad, xr.grad)] + [(k, p.grad, q.grad) for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters())]
for name, a, b in rows:
    s = b.abs().max().item()
    e = (a - b).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item() if s > 0 else float("nan")
    print("%-34s max|ref| %.3e  max err %.3e  rel %.3f  cos %.5f  |a|max %.3e" % (name, s, e, e / max(s, 1e-30), cos, a.abs().max().item()))

# ---- per-layer gradient w.r.t. the conv outputs (pre-norm), needs NHVR_DEBUG_KEEP=1
if os.environ.get("NHVR_DEBUG_KEEP"):
    import torch.nn as nn
    from nhvr_b200 import ops
    convs = [m for m in ref.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    outs = []
    def _hook(mod, inp, out):
        out.retain_grad()
        outs.append(out)
    hooks = [m.register_forward_hook(_hook) for m in convs]
    xr2 = x.detach().clone().requires_grad_(True)
    (ref(xr2) * wgt).mean().backward()
    eng = [e for k, v in net._engines.items() if isinstance(v, list) for e in v][0]
    B, S = eng._bwd, eng._last_S
    for i, o in enumerate(outs):
        gb = B["G"][i]
        d = gb.desc
        got = ops.unpack_nchw(gb, o.shape[1]) / S
        gr = o.grad
        if got.shape != gr.shape:
            print("layer", i, "shape mismatch", tuple(got.shape), tuple(gr.shape)); continue
        s_ = gr.abs().max().item()
        print("G[%d] %-28s max|ref| %.3e rel err %.3f cos %.5f" % (i, tuple(gr.shape), s_, (got - gr).abs().max().item() / max(s_, 1e-30),
              torch.nn.functional.cosine_similarity(got.flatten().double(), gr.flatten().double(), dim=0).item()))
