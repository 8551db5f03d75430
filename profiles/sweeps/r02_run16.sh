#!/bin/bash
# round 2, GPU call 16: device timeline of the stretched whole-stage call (OHP_STRETCH_TRACE): do the walks run beside ramp_convert_kernel?
set -x
O=gpurun_out
for k in 4 8; do
  OHP_STRETCH_TRACE=1 OHP_STRETCHES=$k timeout 300 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 3 --warmup 6 > $O/r02_b16_config2_$k.json 2> $O/r02_b16_config2_$k.err
  grep 'stretch trace' $O/r02_b16_config2_$k.err | tail -$((k*6+1))
done
