#!/bin/bash
# round 2, GPU call 2: the wide any-alignment transform + cut-to-alignment store; consumer warps x groups per step sweep
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t2.log 2>&1; tail -5 $O/r02_t2.log
timeout 300 python profiles/parity_fuzz.py 120 > $O/r02_parity_fuzz_wide.json 2> $O/r02_parity_fuzz_wide.err; tail -c 300 $O/r02_parity_fuzz_wide.json
for wl in config2 config3 config4 config5 mixed; do
  for lib in r1 w8 w10 w12 g2 w10g2 w12g2; do
    export OHP_LIB_CUDA=$PWD/build/libohp_$lib.so
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 10 --warmup 5 > $O/r02_b2_${wl}_$lib.json 2> $O/r02_b2_${wl}_$lib.err
    python - >> $O/r02_sweep2.log <<P
import json
try:
    d=json.load(open("$O/r02_b2_${wl}_$lib.json"))
    print("%-8s %-6s ms %.3f frac %.4f cap %s xsum %s" % ("$wl", "$lib", d["ms_per_step"], d["roofline"]["frac"], d["config"].get("inflight_chunks_per_cta"), d.get("checksum_of_checksums")))
except Exception as e:
    print("$wl $lib FAILED", e)
P
  done
done
unset OHP_LIB_CUDA
cat $O/r02_sweep2.log
for wl in config4 mixed; do
  OHP_LIB_CUDA=$PWD/build/libohp_prof.so timeout 200 python profiles/wait_profile.py $wl > $O/r02_wait2_$wl.log 2>&1; cat $O/r02_wait2_$wl.log
done
