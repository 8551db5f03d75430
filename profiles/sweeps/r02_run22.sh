#!/bin/bash
# round 2, GPU call 22: ncu --set full of the one-walk schedule kernel (EMIT into regions) on configs[1], after the lean recurrence
set -x
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schedule_kernel --launch-skip 9 -c 1 -o $O/r02_prof22_schedule python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 4 --warmup 4 > $O/r02_ncu22.log 2>&1; tail -3 $O/r02_ncu22.log | cut -c1-200
