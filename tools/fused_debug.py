"""Which layers of the two generators take the fused conv + InstanceNorm epilogue (NHVR_DEBUG_FUSED=1 prints the residency numbers)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from nhvr_b200.pipeline import RenderPipeline
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "split3"
pipe = RenderPipeline(size=512, uv_precision=prec, g_precision=prec).to("cuda").eval()
g = pipe.step_graph(B, 512, 512, use_graph=False)
for name in ("engT", "engG"):
    eng = getattr(g, name, None)
    if eng is not None:
        print(name, [(pl.label, f) for pl, f in zip(eng.plans, eng.fused)])
