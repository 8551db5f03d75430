#!/bin/bash
# round 2, GPU call 24: FP64 ramp step for thread-per-stream walks too; the lean loop unrolled 4x / 8x (A/B); ncu of the one-walk kernel
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for v in default unroll4 unroll8; do
  if [ $v = default ]; then unset OHP_LIB_CUDA; else export OHP_LIB_CUDA=$PWD/build/libohp_$v.so; fi
  for wl in config2 config3 config4 config5; do
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --no-check --steps 10 --warmup 6 > $O/r02_b24_${wl}_$v.json 2> $O/r02_b24_${wl}_$v.err
  done
done
unset OHP_LIB_CUDA
python - <<P
import json
for v in ("default","unroll4","unroll8"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b24_%s_%s.json"%(wl,v)))
        print(v,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s | two-pass build %.2f ms"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"],1e3*d["config"]["device_schedule_build_s"]))
    except Exception as e: print(v,wl,"FAILED",e)
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schedule_kernel --launch-skip 9 -c 1 -o $O/r02_prof24_schedule python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 4 --warmup 4 > $O/r02_ncu24.log 2>&1; tail -2 $O/r02_ncu24.log | cut -c1-200
timeout 600 python profiles/parity_fuzz.py 100 > $O/r02_parity_fuzz24.json 2> $O/r02_parity_fuzz24.err; tail -c 300 $O/r02_parity_fuzz24.json
