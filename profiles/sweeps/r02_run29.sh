#!/bin/bash
# round 2, GPU call 29: the one-walk path launched ahead of the host's round trip (default) against the previous form
# (OHP_SLICE_CHUNKS huge: walk on the schedule stream after the host has the regions' offsets), three times each
set -x
O=gpurun_out
for rep in 1 2 3; do
  for mode in ahead after; do
    if [ $mode = after ]; then export OHP_SLICE_CHUNKS=4000000000; else unset OHP_SLICE_CHUNKS; fi
    for wl in config2 config3 config4 config5; do
      timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --no-check --steps 10 --warmup 6 > $O/r02_b29_${wl}_${mode}_$rep.json 2> $O/r02_b29_${wl}_${mode}_$rep.err
    done
  done
done
unset OHP_SLICE_CHUNKS
python - <<P
import json
for wl in ("config2","config3","config4","config5"):
  for mode in ("ahead","after"):
    v=[]
    for rep in (1,2,3):
        try: v.append(json.load(open("$O/r02_b29_%s_%s_%d.json"%(wl,mode,rep)))["value_from_specs"]["ms_per_step"])
        except Exception as e: v.append(float("nan"))
    print(wl,mode," ".join("%.3f"%x for x in v))
P
