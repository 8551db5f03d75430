#!/bin/bash
# round 2, GPU call 17: the walk capped at 168 registers (one of its CTAs fits beside two of ramp_convert_kernel's): timeline, then
# stretch counts against the one-walk path on every config
set -x
O=gpurun_out
OHP_STRETCH_TRACE=1 OHP_STRETCHES=8 timeout 300 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 3 --warmup 6 > $O/r02_b17_trace.json 2> $O/r02_b17_trace.err
grep 'stretch trace' $O/r02_b17_trace.err | tail -49
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for k in 8 4 12 16 0; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b17_${wl}_$k.json 2> $O/r02_b17_${wl}_$k.err
  done
done
unset OHP_STRETCHES
python - <<P
import json
for k in ("8","4","12","16","0"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b17_%s_%s.json"%(wl,k)))
        print("variant",k,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: pass
P
