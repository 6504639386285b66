#!/bin/bash
# GPU-box pass at the end of round 2 (late): ncu launch list of the driver's bench command, full captures of the two K1d
# kernels and of the fast cluster form of the red-black solver, summarised on the box (the reports themselves are too large
# to travel back).
OUT=gpurun_out/${1:-r3b}
mkdir -p $OUT
B="python bench.py --steps 20 --warmup 5 --settle 64 --no-e2e --no-cpu-baseline --no-device-runs --no-dropin-loop"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/bench_launches.csv $B > $OUT/bench_launches.log 2>&1
python tools/launch_summary.py $OUT/bench_launches.csv > $OUT/bench_launches.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkFlightKernel -s 2 -c 1 -f -o $OUT/flight $B > $OUT/flight.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bulkEventKernel -s 2 -c 1 -f -o $OUT/event $B > $OUT/event.log 2>&1
python tools/ncu_summary.py $OUT/flight.ncu-rep $OUT/flight_spl20.txt --units 2000000000 --note "K1d bulkFlightKernel<4, AXIS 0> (flightCoreAxis: 24 FP64 per particle-step), 1e8 electrons, 20 time steps per launch, the driver's bench command"
python tools/ncu_summary.py $OUT/event.ncu-rep $OUT/event_spl20.txt --units 2000000000 --note "K1d bulkEventKernel<PHILOX,false>, claims of 1024 particles, same launch pair"
M="viennaemc_b200/bin/mosfet2D --seed 5 --progress 100000 --steps 120 --transient 40 --avg 40 --red-black 1 --prefix /tmp/m"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sorRedBlackClusterFastKernel --launch-skip 80 --launch-count 1 -f -o $OUT/sor_fast $M > $OUT/sor_fast.log 2>&1
python tools/ncu_summary.py $OUT/sor_fast.ncu-rep $OUT/sor_fast_mosfet.txt --units 0 --note "sorRedBlackClusterFastKernel<512> (16-CTA cluster, DSMEM halo and stopping flags), one non-equilibrium solve of mosfet2D (126x101 grid)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/mosfet_launches.csv $M > /dev/null 2>&1
python tools/launch_summary.py $OUT/mosfet_launches.csv > $OUT/mosfet_launches.txt
R="viennaemc_b200/bin/resistor2D --seed 5 --progress 100000 --steps 200 --transient 50 --avg 50 --red-black 1 --prefix /tmp/r"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/resistor_launches.csv $R > /dev/null 2>&1
python tools/launch_summary.py $OUT/resistor_launches.csv > $OUT/resistor_launches.txt
rm -f $OUT/*.ncu-rep *.txt
ls -la $OUT
