#!/bin/bash
# round 2, GPU call 21: the round's final build -- whole GPU suite, smoke, the default bench command, the reference arm, the launch list
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke21.log 2>&1; tail -2 $O/r02_smoke21.log
timeout 1200 python bench.py > $O/r02_bench21.json 2> $O/r02_bench21.err; tail -c 300 $O/r02_bench21.err; head -c 600 $O/r02_bench21.json
timeout 1200 python bench.py --impl reference > $O/r02_bench21_ref.json 2> $O/r02_bench21_ref.err; head -c 400 $O/r02_bench21_ref.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_bench21_launches.csv python bench.py --steps 2 --warmup 4 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu21.log 2>&1; tail -2 $O/r02_ncu21.log | cut -c1-300
timeout 600 python profiles/parity_fuzz.py 120 > $O/r02_parity_fuzz21.json 2> $O/r02_parity_fuzz21.err; tail -c 300 $O/r02_parity_fuzz21.json
