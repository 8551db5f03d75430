#!/bin/bash
# round 2, GPU call 3: the reworked bench (all five configs, checks against the reference) + the variant sweep again
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t3.log 2>&1; tail -5 $O/r02_t3.log
timeout 900 python bench.py --steps 10 --warmup 5 > $O/r02_bench3.json 2> $O/r02_bench3.err; tail -c 600 $O/r02_bench3.err; head -c 3000 $O/r02_bench3.json
rm -f $O/r02_sweep3.log
for wl in config2 config3 config4 config5 mixed; do
  for lib in w8 w10 w12 g2 w12g2 d0 d1 w12d1; do
    export OHP_LIB_CUDA=$PWD/build/libohp_$lib.so
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b3_${wl}_$lib.json 2> $O/r02_b3_${wl}_$lib.err
    python - >> $O/r02_sweep3.log <<P
import json
try:
    d=json.load(open("$O/r02_b3_${wl}_$lib.json"))
    print("%-8s %-6s ms %.3f frac %.4f cap %s fromspecs %.3f xsum %s" % ("$wl", "$lib", d["ms_per_step"], d["roofline"]["frac"], d["config"].get("inflight_chunks_per_cta"), d["value_from_specs"]["ms_per_step"], d.get("checksum_of_checksums")))
except Exception as e:
    print("$wl $lib FAILED", e)
P
  done
done
unset OHP_LIB_CUDA
cat $O/r02_sweep3.log
