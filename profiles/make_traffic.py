#!/usr/bin/env python
"""profiles/traffic.json from the DRAM counters of one ramp_convert_kernel launch per config:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ramp_convert \\
        --launch-skip 6 -c 1 --csv --log-file gpurun_out/r02_traffic_<cfg>.csv python bench.py --workload <cfg> ... > gpurun_out/r02_traffic_<cfg>.json

(profiles/sweeps/r02_run7.sh).  bench.py scales capture_dram_bytes / capture_algorithmic_bytes to the launch it timed.
    python profiles/make_traffic.py gpurun_out r02"""
import csv
import json
import os
import sys

src, tag = sys.argv[1], sys.argv[2]
out = {}
for cfg in ("config2", "config3", "config4", "config5"):
    c = os.path.join(src, "%s_traffic_%s.csv" % (tag, cfg))
    j = os.path.join(src, "%s_traffic_%s.json" % (tag, cfg))
    if not (os.path.exists(c) and os.path.exists(j)):
        continue
    rows = [r for r in csv.reader(open(c)) if len(r) > 10 and r[0].isdigit()]
    vals = {r[-3]: float(r[-1]) for r in rows}
    line = json.load(open(j))
    out[cfg] = {"capture": "%s: one launch of the bench's own workload (%s), ncu metrics pass, launch-skip 6" % (os.path.basename(c), line["config"]["workload"]),
                "capture_dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
                "capture_dram_read_bytes": vals["dram__bytes_read.sum"], "capture_dram_write_bytes": vals["dram__bytes_write.sum"],
                "capture_algorithmic_bytes": line["roofline"]["algorithmic_bytes_per_launch"],
                "capture_kernel_ms_under_ncu": vals["gpu__time_duration.sum"] / 1e6}
    out[cfg]["ratio"] = out[cfg]["capture_dram_bytes"] / out[cfg]["capture_algorithmic_bytes"]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(json.dumps({k: round(v["ratio"], 4) for k, v in out.items()}))
