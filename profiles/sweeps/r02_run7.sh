#!/bin/bash
# round 2, GPU call 7: one-walk whole stage (value_from_specs), launch list of the default bench command, DRAM traffic of every config
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t7.log 2>&1; tail -5 $O/r02_t7.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench7.json 2> $O/r02_bench7.err; tail -c 600 $O/r02_bench7.err
for wl in config2 config3 config4 config5; do
  OHP_ONE_WALK=0 timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b7_${wl}_twopass.json 2> $O/r02_b7_${wl}_twopass.err
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b7_${wl}_onewalk.json 2> $O/r02_b7_${wl}_onewalk.err
done
python - <<P
import json
for wl in ("config2","config3","config4","config5"):
    for m in ("twopass","onewalk"):
        try:
            d=json.load(open("$O/r02_b7_%s_%s.json"%(wl,m)))
            print(wl,m,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
        except Exception as e: print(wl,m,"FAILED",e)
P
# the launch list of the default command (cold-cache, serialised: shares, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 3 --warmup 4 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu_l7.log 2>&1; tail -2 $O/r02_ncu_l7.log
# DRAM traffic of one launch per config
for wl in config2 config3 config4 config5; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ramp_convert --launch-skip 6 -c 1 --csv --log-file $O/r02_traffic_$wl.csv python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 4 --warmup 4 > $O/r02_traffic_$wl.json 2> $O/r02_traffic_$wl.err
  tail -3 $O/r02_traffic_$wl.csv
done
