#!/bin/bash
# compute-sanitizer passes over the kernels with hand-rolled synchronisation: K1a (TMA pipeline, mbarriers, event queue),
# K1c (CTA-wide event queue with slot handshakes), K1d (warp-private lists), the SOR kernels (cluster barriers + DSMEM).
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
T1=(tests/test_bulk_gpu.py -k "philox_against_oracle or (replay and mixed and spl1)")
T2=(tests/test_device_gpu.py -k sor_variants)
# later additions of round 2: grain clock in K1d and in the look-ahead copy, single-layer valleys / samplers (K1b, K1c)
T3=(tests/test_sl_gpu.py tests/test_lookahead_gpu.py -k "philox_against_oracle or grain_clocks")
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 $PY "${T3[@]}" > $OUT/${tool}_grain_sl.log 2>&1
  echo "$tool grain+sl: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${tool}_grain_sl.log | tr '\n' ' ')"
  [ -n "$2" ] && continue
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 $PY "${T1[@]}" > $OUT/${tool}_bulk.log 2>&1
  echo "$tool bulk: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${tool}_bulk.log | tr '\n' ' ')"
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 $PY "${T2[@]}" > $OUT/${tool}_sor.log 2>&1
  echo "$tool sor: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${tool}_sor.log | tr '\n' ' ')"
done
