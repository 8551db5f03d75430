#!/bin/bash
# round 2, GPU call 19: testing the explanation of call 18 (sub-partition register files keep the walk's CTAs out): a walk
# squeezed to 80 registers, and ramp_convert_kernel with 8 warps per CTA (7 consumers) -- do the walk's CTAs start beside it now?
# Then the whole GPU suite on the default build (one walk into bounded regions).
set -x
O=gpurun_out
for v in sched80 w7; do
  export OHP_LIB_CUDA=$PWD/build/libohp_$v.so
  OHP_STRETCH_TRACE=1 OHP_STRETCHES=8 timeout 300 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 2 --warmup 6 > $O/r02_b19_trace_$v.json 2> $O/r02_b19_trace_$v.err
  grep 'stretch trace' $O/r02_b19_trace_$v.err | tail -114 | grep -v 'CTA' | head -24 | cut -c1-100
  grep 'stretch trace' $O/r02_b19_trace_$v.err | tail -66 | grep 'CTA' | awk 'NR%6==1' | cut -c1-130
  for k in 8 4 0; do
    OHP_STRETCHES=$k timeout 300 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 10 --warmup 6 > $O/r02_b19_${v}_$k.json 2> $O/r02_b19_${v}_$k.err
  done
done
unset OHP_LIB_CUDA
python - <<P
import json
for v in ("sched80","w7"):
  for k in (8,4,0):
    try:
        d=json.load(open("$O/r02_b19_%s_%d.json"%(v,k)))
        print("variant",v,"stretches",k,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(v,k,"FAILED",e)
P
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
