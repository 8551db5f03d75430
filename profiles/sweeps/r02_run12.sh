#!/bin/bash
# round 2, GPU call 12: stretches again with a 256-thread scan kernel (the 1024-thread one could not run beside
# ramp_convert_kernel's persistent CTAs and serialised the pipeline); two units per lane in the compact transform (A/B)
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for k in 5 3 0; do
  for wl in config2 config3 config4 config5; do
    export OHP_STRETCHES=$k
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 6 > $O/r02_b12_${wl}_$k.json 2> $O/r02_b12_${wl}_$k.err
  done
done
unset OHP_STRETCHES
export OHP_LIB_CUDA=$PWD/build/libohp_any2.so
for wl in config3 config4 mixed; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 10 --warmup 6 > $O/r02_b12_${wl}_any2.json 2> $O/r02_b12_${wl}_any2.err
done
unset OHP_LIB_CUDA
timeout 300 python bench.py --workload mixed --no-e2e --no-cpu-baseline --steps 10 --warmup 6 > $O/r02_b12_mixed_5.json 2> $O/r02_b12_mixed_5.err
python - <<P
import json
for k in ("5","3","0","any2"):
  for wl in ("config2","config3","config4","config5","mixed"):
    try:
        d=json.load(open("$O/r02_b12_%s_%s.json"%(wl,k)))
        print("variant",k,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s | exact %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"],d.get("bit_exact")))
    except Exception as e: pass
P
