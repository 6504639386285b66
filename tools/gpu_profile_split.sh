#!/bin/bash
# GPU-box pass: ncu launch list and full captures of the K1d kernels (flight + event) on the bench-sized ensemble.
OUT=gpurun_out/r2b
mkdir -p $OUT
S="python tools/kernel_sweep.py --steps 32 --settle 64 --configs 3:16:4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $S > $OUT/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkFlightKernel -s 5 -c 1 -f -o $OUT/flight $S > $OUT/flight.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkEventKernel -s 5 -c 1 -f -o $OUT/event $S > $OUT/event.log 2>&1
ls -la $OUT
