#!/bin/bash
# last-session evidence bundle (run under gpurun from the repo root): backward tests, training profiles, ncu of the lean sampler,
# launch list and event-bracketed profile of one strict frame step
python -m pytest tests -m gpu -x -q -k "backward or training or pretrain" 2>&1 | tail -1
python tools/train_profile.py e2e > gpurun_out/r02c_train_profile_e2e_gx.log 2>&1
python tools/train_profile.py uv > gpurun_out/r02c_train_profile_uv_gx.log 2>&1
grep -h "in_bwd\|sum of" gpurun_out/r02c_train_profile_*_gx.log
NCU="ncu --profile-from-start off --clock-control none"
$NCU --set full -k regex:texture_sample --launch-count 1 -o gpurun_out/r02c_ncu_sampler_lean_full -f python tools/ncu_step.py strict > /dev/null 2>&1
ncu -i gpurun_out/r02c_ncu_sampler_lean_full.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_sampler_lean_raw.csv 2>/dev/null
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02c_ncu_launches_strict.csv python tools/ncu_step.py strict > /dev/null 2>&1
python tools/step_profile.py 8 split3 512 split3 > gpurun_out/r02c_step_profile_strict_c8.log 2>&1
tail -2 gpurun_out/r02c_step_profile_strict_c8.log; grep sampler gpurun_out/r02c_step_profile_strict_c8.log
