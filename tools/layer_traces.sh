# per-CTA cycle traces of the non-ResnetBlock layers of the strict frame step (tools/conv_trace.py)
run() { echo "== $*"; timeout 120 python tools/conv_trace.py "$@" 2>&1 | grep -E "us  |conv trace\]|kcp" | cut -c1-360; }
run CONV_TRANSPOSE 128 64 3 2 1 8 256 256 Z RAW_STATS split3
run CONV_TRANSPOSE 256 128 3 2 1 8 128 128 Z RAW_STATS split3
run CONV_TRANSPOSE 96 48 3 2 1 8 256 256 Z RAW_STATS split3
run CONV 12 48 7 1 3 8 512 512 R RAW_STATS split3
run CONV 48 4 7 1 3 8 512 512 R BIAS_ACT_F32 split3
run CONV 96 192 3 2 1 8 256 256 Z RAW_STATS split3
