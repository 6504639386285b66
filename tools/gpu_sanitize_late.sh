OUT=gpurun_out/r2z_sanitize; mkdir -p $OUT
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 $PY tests/test_device_gpu.py -k "sor_variants or particle_step_replays" > $OUT/${tool}_device.log 2>&1
  echo "$tool device: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${tool}_device.log | tr '\n' ' ')"
  timeout 600 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 $PY tests/test_bulk_gpu.py -k "philox_against_oracle" > $OUT/${tool}_bulk.log 2>&1
  echo "$tool bulk: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${tool}_bulk.log | tr '\n' ' ')"
done
