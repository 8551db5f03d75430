#!/bin/bash
# round 2, GPU call 10: lean ramp recurrence in the bulk step (reciprocals of a run's divisors taken side by side) and
# thread-per-stream against warp-per-stream walks where streams are many (configs[2], configs[3])
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py -m gpu -x -q 2>&1 | tail -3
for team in auto 1 32; do
  for wl in config2 config3 config4 config5; do
    if [ $team = auto ]; then unset OHP_SCHED_TEAM; else export OHP_SCHED_TEAM=$team; fi
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b10_${wl}_$team.json 2> $O/r02_b10_${wl}_$team.err
  done
done
unset OHP_SCHED_TEAM
python - <<P
import json
for team in ("auto","1","32"):
  for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b10_%s_%s.json"%(wl,team)))
        print("team",team,wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s | two-pass build %.2f ms"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"],1e3*d["config"]["device_schedule_build_s"]))
    except Exception as e: print(team,wl,"FAILED",e)
P
timeout 600 python profiles/parity_fuzz.py 100 > $O/r02_parity_fuzz10.json 2> $O/r02_parity_fuzz10.err; tail -c 300 $O/r02_parity_fuzz10.json
