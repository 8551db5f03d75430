#!/bin/bash
# round 2, GPU call 11: ohp_run_streams_device walking in stretches beside ramp_convert_kernel; stretch counts; the whole
# GPU suite on the build
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -5
for k in 5 0 3 4 6 8; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b11_${wl}_$k.json 2> $O/r02_b11_${wl}_$k.err
  done
done
unset OHP_STRETCHES
python - <<P
import json
for k in (5,0,3,4,6,8):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b11_%s_%d.json"%(wl,k)))
        print("stretches",k,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(k,wl,"FAILED",e)
P
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python profiles/parity_fuzz.py 100 > $O/r02_parity_fuzz11.json 2> $O/r02_parity_fuzz11.err; tail -c 300 $O/r02_parity_fuzz11.json
