#!/bin/bash
# GPU-box pass of round 2: the driver's bench command, its ncu launch list, full captures of the two K1d kernels.
OUT=gpurun_out/${1:-r2g}
mkdir -p $OUT
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -3 $OUT/bench_n1.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ceiling frac", d["e2e"].get("copy_ceiling", {}).get("frac"), "long", d["e2e"].get("long_run", {}).get("value"))
print("roofline frac", d["roofline"]["frac"], "flight ms", d["roofline"]["avg_launch_ms"], "event ms", d["roofline"]["event_kernel"]["avg_launch_ms"])
print("dropin", d.get("dropin_loop"))
print("one step", d["one_step_per_launch"]["value"], "cpu", d.get("cpu_baseline"))
PY
if [ "$2" != "noprof" ]; then
B="python bench.py --steps 20 --warmup 5 --settle 64 --no-e2e --no-cpu-baseline --no-device-runs --no-dropin-loop"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/bench_launches.csv $B > $OUT/bench_launches.log 2>&1
python tools/launch_summary.py $OUT/bench_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkFlightKernel -s 2 -c 1 -f -o $OUT/flight $B > $OUT/flight.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkEventKernel -s 2 -c 1 -f -o $OUT/event $B > $OUT/event.log 2>&1
fi
ls -la $OUT
