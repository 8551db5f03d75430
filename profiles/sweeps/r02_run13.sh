#!/bin/bash
# round 2, GPU call 13: equal stretches (4 / 8 / 12 / 16) against the one-walk path; a 96-register walk (7 warps per SM fit
# beside ramp_convert_kernel instead of 4)
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for k in 8 4 12 16 0; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b13_${wl}_$k.json 2> $O/r02_b13_${wl}_$k.err
  done
done
export OHP_LIB_CUDA=$PWD/build/libohp_sched96.so
for k in 8 0; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b13_${wl}_r96_$k.json 2> $O/r02_b13_${wl}_r96_$k.err
  done
done
unset OHP_LIB_CUDA OHP_STRETCHES
python - <<P
import json
for k in ("8","4","12","16","0","r96_8","r96_0"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b13_%s_%s.json"%(wl,k)))
        print("variant",k,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(k,wl,"FAILED",e)
P
