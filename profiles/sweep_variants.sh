#!/bin/bash
# usage: profiles/sweep_variants.sh <bench args...>   -- runs bench.py once per build/libohp_*.so (kernel tunable experiments)
for lib in build/libohp_*.so; do
  name=$(basename $lib .so)
  OHP_LIB_CUDA=$PWD/$lib python bench.py --no-e2e --no-cpu-baseline "$@" 2>/tmp/sweep_err.txt | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('%-24s ms/step %.3f  achieved %.0f GB/s  frac %.3f  xsum %s' % ('$name', d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d.get('checksum_of_checksums')))
except Exception:
    print('%-24s FAILED: %s' % ('$name', open('/tmp/sweep_err.txt').read().strip().splitlines()[-1][:160] if open('/tmp/sweep_err.txt').read().strip() else 'no output'))"
done
