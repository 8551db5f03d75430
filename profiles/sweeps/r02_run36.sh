#!/bin/bash
# round 2, GPU call 36: the final build once more -- GPU suite, smoke, the default bench command and the reference arm (the lines kept
# under profiles/), the launch list, and a long run of the GPU-vs-oracle campaign
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke36.log 2>&1; tail -1 $O/r02_smoke36.log
timeout 1200 python bench.py --impl reference > $O/r02_bench36_ref.json 2> $O/r02_bench36_ref.err; head -c 200 $O/r02_bench36_ref.json
timeout 1200 python bench.py > $O/r02_bench36.json 2> $O/r02_bench36.err; tail -c 300 $O/r02_bench36.err; head -c 200 $O/r02_bench36.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_bench36_launches.csv python bench.py --steps 2 --warmup 4 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu36.log 2>&1; tail -1 $O/r02_ncu36.log | cut -c1-200
timeout 700 python profiles/parity_fuzz.py 560 > $O/r02_parity_fuzz36.json 2> $O/r02_parity_fuzz36.err; tail -c 300 $O/r02_parity_fuzz36.json
