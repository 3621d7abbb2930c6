# fused conv + InstanceNorm vs conv + apply on the three ResnetBlock layer shapes (tools/fused_trace.py), with the per-CTA trace
for cfg in "256 8 f16" "256 8 split3" "192 8 f16" "192 8 split3" "256 2 f16"; do
  timeout 100 python tools/fused_trace.py $cfg 2>&1 | grep -v "^plan"
  NHVR_CONV_TRACE=1 timeout 100 python tools/fused_trace.py $cfg 2>&1 | grep "trace fused" | tail -1
done
