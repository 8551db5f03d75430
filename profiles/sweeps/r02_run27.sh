#!/bin/bash
# round 2, GPU call 27: the FP64 recurrence at 31 instructions per message (the end of a lane's message taken from the next lane after the loop, lane number held in a register)
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py tests/test_container.py -m gpu -x -q 2>&1 | tail -3
for wl in config2 config3 config4 config5; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-configs --steps 10 --warmup 6 > $O/r02_b27_$wl.json 2> $O/r02_b27_$wl.err
done
python - <<P
import json
for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b27_%s.json"%wl))
        print(wl,"kernel %.3f ms frac %.4f exact %s | from specs %.3f ms frac %.4f same %s | two-pass build %.2f ms"%(d["ms_per_step"],d["roofline"]["frac"],d.get("bit_exact"),d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"],1e3*d["config"]["device_schedule_build_s"]))
    except Exception as e: print(wl,"FAILED",e)
P
timeout 600 python profiles/parity_fuzz.py 150 > $O/r02_parity_fuzz27.json 2> $O/r02_parity_fuzz27.err; tail -c 300 $O/r02_parity_fuzz27.json
