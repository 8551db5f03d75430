#!/bin/bash
# round 2, GPU call 37: the lean transform for chunks that start mid-word on a frame boundary of the 16-byte grid (16-bit stereo
# behind a split, ...): parity suites, fuzz, every config
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python profiles/parity_fuzz.py 200 > $O/r02_parity_fuzz37.json 2> $O/r02_parity_fuzz37.err; tail -c 250 $O/r02_parity_fuzz37.json
for wl in config2 config3 config4 config5 mixed; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --steps 10 --warmup 6 > $O/r02_b37_$wl.json 2> $O/r02_b37_$wl.err
done
python - <<P
import json
for wl in ("config2","config3","config4","config5","mixed"):
    try:
        d=json.load(open("$O/r02_b37_%s.json"%wl))
        print(wl,"kernel %.3f ms frac %.4f exact %s cap %s | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d.get("bit_exact"),d["config"].get("inflight_chunks_per_cta"),d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(wl,"FAILED",e)
P
