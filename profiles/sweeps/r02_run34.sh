#!/bin/bash
# round 2, GPU calls 34 / 35: the bench as the driver launches it at N = 2 (and N = 8): final build
set -x
O=gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 5 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err; tail -c 500 $O/r02_bench_${N}gpu.err; head -c 400 $O/r02_bench_${N}gpu.json
python - <<P
import json
d=json.load(open("$O/r02_bench_${N}gpu.json"))
print("value %.4g ms %.3f frac %.4f exact %s n_gpus %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["bit_exact"],d["n_gpus"]))
e=d["e2e"]; print("e2e %.4g %s ms %.1f pcie %.1f ceiling %.1f frac %.3f exact %s"%(e["value"],e["unit"],e["ms_per_step"],e["pcie_gbs"],e["ceiling_gbs"],e["frac_of_ceiling"],e["bit_exact"]))
for c in d["configs"]: print(c["workload"][:30],"%.3f ms frac %.4f exact %s n=%s %s"%(c["ms_per_launch"],c["frac"],c["bit_exact"],c["streams_checked"],c["scaling"]))
P
