#!/bin/bash
# round 2, GPU call 9: the schedule walk's ramp recurrence with precomputed reciprocals (A/B against the previous build),
# the GPU schedule suites on it, and one ncu --set full capture of schedule_kernel<EMIT, 32> on configs[1]
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for lib in olddiv new; do
  for wl in config2 config3 config4 config5; do
    if [ $lib = olddiv ]; then export OHP_LIB_CUDA=$PWD/build/libohp_olddiv.so; else unset OHP_LIB_CUDA; fi
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b9_${wl}_$lib.json 2> $O/r02_b9_${wl}_$lib.err
  done
done
unset OHP_LIB_CUDA
python - <<P
import json
for lib in ("olddiv","new"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b9_%s_%s.json"%(wl,lib)))
        print(lib,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(lib,wl,"FAILED",e)
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schedule_kernel --launch-skip 4 -c 1 -o $O/r02_prof9_schedule python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --steps 4 --warmup 4 > $O/r02_ncu9.log 2>&1; tail -3 $O/r02_ncu9.log
timeout 600 python profiles/parity_fuzz.py 150 > $O/r02_parity_fuzz9.json 2> $O/r02_parity_fuzz9.err; tail -c 400 $O/r02_parity_fuzz9.json
