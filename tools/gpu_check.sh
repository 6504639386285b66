#!/bin/bash
# One GPU-box pass used during development: host-run tests, the bench line, ncu launch list + full capture of K1c.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bulk_gpu.py tests/test_zz_host_run_gpu.py -m gpu -x -q -k "host_resident or sharding or large" > gpurun_out/host_run_tests.log 2>&1
tail -15 gpurun_out/host_run_tests.log
timeout 400 python bench.py --no-device-runs --no-cpu-baseline > gpurun_out/bench_n1_hostrun.json 2> gpurun_out/bench_n1_hostrun.err
tail -3 gpurun_out/bench_n1_hostrun.err
python -c '
import json
d = json.loads(open("gpurun_out/bench_n1_hostrun.json").read().strip().splitlines()[-1])
print(d["value"], json.dumps(d["e2e"]))'
B="python bench.py --steps 64 --warmup 8 --settle 64 --no-e2e --no-cpu-baseline --no-device-runs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_q_launches.csv $B > gpurun_out/r1_q_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bulkDeferKernel -s 4 -c 1 -f -o gpurun_out/r1_q_defer $B > gpurun_out/r1_q_full.log 2>&1
ls -la gpurun_out | grep r1_q
