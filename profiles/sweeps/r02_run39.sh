#!/bin/bash
# round 2, GPU call 39: the parity campaign on new seeds, now through both whole-stage calls (host buffers and HBM-resident)
set -x
O=gpurun_out
timeout 700 python profiles/parity_fuzz.py 540 100000 > $O/r02_parity_fuzz39.json 2> $O/r02_parity_fuzz39.err; tail -c 600 $O/r02_parity_fuzz39.json; tail -3 $O/r02_parity_fuzz39.err
