#!/bin/bash
# round 2, GPU call 1: OHP_DEFER_RELEASE as the default build -- every GPU suite, the fuzz campaign, and bench lines of
# every BASELINE config next to the round-1 build (build/libohp_nodefer.so)
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t1.log 2>&1; tail -3 $O/r02_t1.log
timeout 400 python profiles/parity_fuzz.py 200 > $O/r02_parity_fuzz_defer.json 2> $O/r02_parity_fuzz_defer.err; tail -c 600 $O/r02_parity_fuzz_defer.json
for wl in config2 config3 config4 config5 mixed; do
  for lib in default nodefer; do
    if [ $lib = default ]; then unset OHP_LIB_CUDA; else export OHP_LIB_CUDA=$PWD/build/libohp_nodefer.so; fi
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 10 --warmup 5 > $O/r02_b1_${wl}_$lib.json 2> $O/r02_b1_${wl}_$lib.err
    python - <<P
import json
try:
    d=json.load(open("$O/r02_b1_${wl}_$lib.json"))
    print("$wl $lib ms %.3f frac %.4f cap %s xsum %s" % (d["ms_per_step"], d["roofline"]["frac"], d["config"].get("inflight_chunks_per_cta"), d.get("checksum_of_checksums")))
except Exception as e:
    print("$wl $lib FAILED", e)
P
  done
done
unset OHP_LIB_CUDA
for wl in config2 config4 config5 mixed; do
  OHP_LIB_CUDA=$PWD/build/libohp_prof.so timeout 200 python profiles/wait_profile.py $wl > $O/r02_wait_$wl.log 2>&1; cat $O/r02_wait_$wl.log
done
