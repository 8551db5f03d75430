#!/bin/bash
# round 2, GPU call 33: streams-per-warp rule on the final build: GPU suite, specs-to-bytes on every config
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in config2 config3 config4 config5; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --steps 10 --warmup 6 > $O/r02_b33_$wl.json 2> $O/r02_b33_$wl.err
done
python - <<P
import json
for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b33_%s.json"%wl))
        print(wl,"kernel %.3f ms frac %.4f exact %s | from specs %.3f ms frac %.4f same %s | two-pass build %.2f ms"%(d["ms_per_step"],d["roofline"]["frac"],d.get("bit_exact"),d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"],1e3*d["config"]["device_schedule_build_s"]))
    except Exception as e: print(wl,"FAILED",e)
P
