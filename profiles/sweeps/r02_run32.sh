#!/bin/bash
# round 2, GPU call 32: thread-per-stream walks with fewer streams per warp (OHP_SCHED_LANES): how much of their time is divergence?
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -2
for lanes in 32 16 8 4 2 1; do
  for wl in config3 config4 config5; do
    OHP_SCHED_LANES=$lanes timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --no-check --steps 10 --warmup 6 > $O/r02_b32_${wl}_$lanes.json 2> $O/r02_b32_${wl}_$lanes.err
  done
done
python - <<P
import json
for wl in ("config3","config4","config5"):
  row=[]
  for lanes in (32,16,8,4,2,1):
    try:
        d=json.load(open("$O/r02_b32_%s_%d.json"%(wl,lanes)))
        row.append("%d: %.3f (%s)"%(lanes,d["value_from_specs"]["ms_per_step"],"same" if d["value_from_specs"]["same_checksums"] else "DIFF"))
    except Exception as e: row.append("%d: FAILED"%lanes)
  print(wl,"kernel alone %.3f |"%d["ms_per_step"]," ".join(row))
P
