#!/bin/bash
# end-of-round verification bundle (run under gpurun from the repo root)
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
(cd tests/native && timeout 200 ./conv_selftest 2>&1 | tail -1 && timeout 200 ./bwd_selftest 2>&1 | tail -1)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][0])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["train"]["steps_per_s"], d["train"]["e2e_step"]["steps_per_s"], d["cpu_baseline"]["value"], {k:(round(v["frac"],3), round(v["ms_per_frame_step"],3)) for k,v in d["roofline_memory_kernels"].items()})
PY
