#!/bin/bash
# round 2, GPU call 20: consumer warps per CTA (10 / 11 / 12 against the shipped 8) on the compact-path kernel, every config
set -x
O=gpurun_out
for v in w8 w10 w11 w12; do
  if [ $v = w8 ]; then unset OHP_LIB_CUDA; else export OHP_LIB_CUDA=$PWD/build/libohp_$v.so; fi
  for wl in config2 config3 config4 config5 mixed; do
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --steps 10 --warmup 6 > $O/r02_b20_${wl}_$v.json 2> $O/r02_b20_${wl}_$v.err
  done
done
unset OHP_LIB_CUDA
python - <<P
import json
for v in ("w8","w10","w11","w12"):
  row=[]
  for wl in ("config2","config3","config4","config5","mixed"):
    try:
        d=json.load(open("$O/r02_b20_%s_%s.json"%(wl,v)))
        row.append("%s %.4f (cap %s, exact %s)"%(wl,d["roofline"]["frac"],d["config"].get("inflight_chunks_per_cta"),d.get("bit_exact")))
    except Exception as e: row.append(wl+" FAILED")
  print("variant",v," | ".join(row))
P
