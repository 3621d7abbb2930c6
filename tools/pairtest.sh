run() { echo "== $* (PAIR=$P)"; NHVR_CONV_PAIR=$P timeout 120 python tools/conv_trace.py "$@" 2>&1 | grep -E "us  |conv trace\]|Error|error" | cut -c1-330; }
for P in "" 2; do
  run CONV 64 73 7 1 3 8 512 512 R BIAS_ACT_F32 split3
  run CONV 64 128 3 2 1 8 512 512 Z RAW_STATS split3
  run CONV 48 96 3 2 1 8 512 512 Z RAW_STATS split3
  run CONV 64 73 7 1 3 8 512 512 R BIAS_ACT_F32
done
