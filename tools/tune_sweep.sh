#!/bin/bash
# sweep ring shapes of the conv kernel on the PERF cases of the native self-test (argument 100 = perf only)
for t in "" "2,3,2,5" "2,3,2,12" "2,3,4,7" "2,4,1,24" "4,2,2,4" "4,2,2,10" "4,3,4,5" "6,2,2,8" "8,2,2,6" "2,2,8,3"; do
  echo "=== NHVR_CONV_TUNE=$t"
  NHVR_CONV_TUNE=$t timeout 120 tests/native/conv_selftest 100 2>&1 | grep PERF | cut -c1-150
done
