#!/bin/bash
# round 2, GPU call 30: the final build -- whole GPU suite, smoke, the default bench command, the reference arm, the launch list,
# ncu --set full of the one-walk schedule kernel
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke30.log 2>&1; tail -2 $O/r02_smoke30.log
timeout 1200 python bench.py > $O/r02_bench30.json 2> $O/r02_bench30.err; tail -c 300 $O/r02_bench30.err; head -c 300 $O/r02_bench30.json
timeout 1200 python bench.py --impl reference > $O/r02_bench30_ref.json 2> $O/r02_bench30_ref.err; head -c 300 $O/r02_bench30_ref.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_bench30_launches.csv python bench.py --steps 2 --warmup 4 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu30.log 2>&1; tail -2 $O/r02_ncu30.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schedule_kernel --launch-skip 9 -c 1 -o $O/r02_prof30_schedule python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --no-configs --steps 4 --warmup 4 > $O/r02_ncu30b.log 2>&1; tail -2 $O/r02_ncu30b.log | cut -c1-200
timeout 600 python profiles/parity_fuzz.py 150 > $O/r02_parity_fuzz30.json 2> $O/r02_parity_fuzz30.err; tail -c 300 $O/r02_parity_fuzz30.json
