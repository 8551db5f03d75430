#!/bin/bash
# usage: profiles/sweep_variants.sh <bench args...>   -- runs bench.py once per build/libohp_*.so (kernel tunable experiments)
for lib in build/libohp_*.so; do
  name=$(basename $lib .so)
  OHP_LIB_CUDA=$PWD/$lib python bench.py --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-24s ms/step %.3f  achieved %.0f GB/s  frac %.3f' % ('$name', d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac']))"
done
