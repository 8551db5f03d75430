#!/bin/bash
# round 2, GPU call 18: when do the CTAs of a walk that is enqueued beside ramp_convert_kernel start, and how long do they run?
set -x
O=gpurun_out
OHP_STRETCH_TRACE=1 OHP_STRETCHES=8 timeout 300 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 2 --warmup 6 > $O/r02_b18_trace.json 2> $O/r02_b18_trace.err
grep 'stretch trace' $O/r02_b18_trace.err | tail -114 | cut -c1-120
