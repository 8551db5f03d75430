#!/bin/bash
# round 2, GPU call 5: compact general path + adaptive routing against the instruction-fetch stalls of the mixed-format config
# (same measurements as call 4)
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t5.log 2>&1; tail -5 $O/r02_t5.log
rm -f $O/r02_sweep5.log
for wl in config4 mixed config2 config3 config5; do
  for lib in w8 noadapt w10 w12; do
    export OHP_LIB_CUDA=$PWD/build/libohp_$lib.so
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b5_${wl}_$lib.json 2> $O/r02_b5_${wl}_$lib.err
    python - >> $O/r02_sweep5.log <<P
import json
try:
    d=json.load(open("$O/r02_b5_${wl}_$lib.json"))
    print("%-8s %-6s ms %.3f frac %.4f cap %s fromspecs %.3f xsum %s" % ("$wl", "$lib", d["ms_per_step"], d["roofline"]["frac"], d["config"].get("inflight_chunks_per_cta"), d["value_from_specs"]["ms_per_step"], d.get("checksum_of_checksums")))
except Exception as e:
    print("$wl $lib FAILED", e)
P
  done
done
unset OHP_LIB_CUDA
cat $O/r02_sweep5.log
OHP_LIB_CUDA=$PWD/build/libohp_prof.so timeout 200 python profiles/wait_profile.py config4 16384 0.25 > $O/r02_wait5_config4.log 2>&1; cat $O/r02_wait4_config4.log
ncu --set full --clock-control none --import-source on -k regex:ramp_convert --launch-skip 6 -c 1 -o $O/r02_prof5_config4 python bench.py --workload config4 --no-e2e --no-cpu-baseline --no-check --steps 4 --warmup 4 > $O/r02_ncu5.log 2>&1; tail -3 $O/r02_ncu4.log
timeout 900 python bench.py --steps 10 --warmup 5 > $O/r02_bench5.json 2> $O/r02_bench5.err; tail -c 600 $O/r02_bench5.err; head -c 6000 $O/r02_bench5.json
