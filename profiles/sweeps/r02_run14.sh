#!/bin/bash
# round 2, GPU call 14: the schedule kernels ask for the same shared-memory split as ramp_convert_kernel (an SM cannot change
# its split while CTAs are resident, so without this their CTAs wait for the persistent kernel to end): do the walks overlap now?
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for k in 8 4 2 0; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b14_${wl}_$k.json 2> $O/r02_b14_${wl}_$k.err
  done
done
export OHP_SCHED_TEAM=1
for k in 8; do
  for wl in config2; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b14_${wl}_t1_$k.json 2> $O/r02_b14_${wl}_t1_$k.err
  done
done
unset OHP_SCHED_TEAM OHP_STRETCHES
python - <<P
import json
for k in ("8","4","2","0","t1_8"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b14_%s_%s.json"%(wl,k)))
        print("variant",k,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: pass
P
