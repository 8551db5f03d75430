#!/bin/bash
# round 2, GPU call 31: what the thread-per-stream walks of configs[2] and configs[3] spend their time on (ncu --set full, source level)
set -x
O=gpurun_out
for wl in config3 config4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:schedule_kernel --launch-skip 9 -c 1 -o $O/r02_prof31_$wl python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --no-configs --steps 4 --warmup 4 > $O/r02_ncu31_$wl.log 2>&1; tail -2 $O/r02_ncu31_$wl.log | cut -c1-200
done
