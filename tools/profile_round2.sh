#!/bin/bash
# Reproduces the ncu evidence under profiles/ in ONE gpurun call (run from the repo root on a B200):
#   r02b_ncu_launches_<mode>.csv   launch list of one frame step (gpu__time_duration per kernel) - strict and fast
#   r02b_conv_dram.csv             dram bytes read / written + duration of every conv launch of the strict step
#   r02b_ncu_uvres_fused_full.ncu-rep    ncu --set full of one split-precision 256->256 ResnetBlock conv (the dominant kernel class)
#   r02b_ncu_gres_fused_full.ncu-rep     same for the 192->192 CTA-pair conv of the temporal generator
set -x
OUT=gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
for mode in strict fast; do
  $NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/r02b_ncu_launches_$mode.csv python tools/ncu_step.py $mode > /dev/null 2>&1
done
$NCU -k regex:conv_shiftgemm --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum \
  --csv --log-file $OUT/r02b_conv_dram.csv python tools/ncu_step.py strict > /dev/null 2>&1
# launch index of the 4th uv res conv (pack, shift, stem, 2 downs come first): kernel-name filter + skip
$NCU --set full --import-source on -k regex:conv_shiftgemm --launch-skip 6 --launch-count 1 -o $OUT/r02b_ncu_uvres_fused_full -f python tools/ncu_step.py strict > /dev/null 2>&1
$NCU --set full --import-source on -k regex:conv_shiftgemm --launch-skip 21 --launch-count 1 -o $OUT/r02b_ncu_gres_fused_full -f python tools/ncu_step.py strict > /dev/null 2>&1
$NCU --set full -k regex:in_apply_rows --launch-skip 1 --launch-count 1 -o $OUT/r02b_ncu_inapply_hilo_full -f python tools/ncu_step.py strict > /dev/null 2>&1
$NCU --set full -k regex:texture_sample --launch-count 1 -o $OUT/r02b_ncu_sampler_full -f python tools/ncu_step.py strict > /dev/null 2>&1
ls -la $OUT/*.ncu-rep $OUT/r02b_*.csv
# per-launch step profiles (event-bracketed, un-graphed) and per-CTA cycle traces (NHVR_CONV_TRACE) behind the tables of DESIGN.md 5.1
python tools/step_profile.py 8 split3 512 split3 > $OUT/r02b_step_profile_strict_c8.log 2>&1
python tools/step_profile.py 8 f16 512 f16 > $OUT/r02b_step_profile_fast_c8.log 2>&1
bash tools/fused_sweep.sh > $OUT/r02b_fused_sweep.log 2>&1
bash tools/layer_traces.sh > $OUT/r02b_layer_traces.log 2>&1
# training: launch lists of one end-to-end step and one UV pre-train step
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/r02b_ncu_launches_train.csv python tools/ncu_train_step.py > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/r02b_ncu_launches_train_uv.csv python tools/ncu_train_step.py uv > /dev/null 2>&1
# summaries are made on the CPU box: python tools/ncu_summarise.py launches|traffic ...
