#!/bin/bash
# round 2, GPU call 8 (2 GPUs): the bench as the driver launches it at N = 2 -- configs[4] sharded 32768 streams per rank,
# every rank's checksums against the reference, the gathered checksum of checksums; gloo-free NCCL plumbing only
set -x
O=gpurun_out
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > $O/r02_bench8_2gpu.json 2> $O/r02_bench8_2gpu.err; tail -c 800 $O/r02_bench8_2gpu.err; head -c 1200 $O/r02_bench8_2gpu.json
timeout 600 python bench.py --workload config2 --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b8_config2.json 2> $O/r02_b8_config2.err
for wl in config3 config4 config5; do timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-check --steps 10 --warmup 5 > $O/r02_b8_$wl.json 2> $O/r02_b8_$wl.err; done
python - <<P
import json
for wl in ("config2","config3","config4","config5"):
    try:
        d=json.load(open("$O/r02_b8_%s.json"%wl))
        print(wl,"kernel %.3f ms frac %.4f | from specs %.3f ms frac %.4f same %s"%(d["ms_per_step"],d["roofline"]["frac"],d["value_from_specs"]["ms_per_step"],d["value_from_specs"]["frac"],d["value_from_specs"]["same_checksums"]))
    except Exception as e: print(wl,"FAILED",e)
P
