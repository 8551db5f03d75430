#!/bin/bash
# round 2, GPU call 6: everything so far on the GPU -- suites (element-level events, P2 modes, host-path sentinels, device-resident
# whole stage), the fuzz campaign with the elements generator, the full bench line
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_t6.log 2>&1; tail -5 $O/r02_t6.log
timeout 400 python profiles/parity_fuzz.py 180 > $O/r02_parity_fuzz.json 2> $O/r02_parity_fuzz.err; tail -c 900 $O/r02_parity_fuzz.json; tail -3 $O/r02_parity_fuzz.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench6.json 2> $O/r02_bench6.err; tail -c 600 $O/r02_bench6.err; head -c 1500 $O/r02_bench6.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/r02_bench6_ref.json 2> $O/r02_bench6_ref.err; head -c 600 $O/r02_bench6_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke6.log 2>&1; tail -2 $O/r02_smoke6.log
