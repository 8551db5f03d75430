"""Parity campaign on the GPU box: many seeds of the randomized workloads through the whole stage on the GPU
(ohp_run_streams_host: device-built descriptors + ramp_convert_kernel) against the C oracle, byte for byte on
everything a chunk covers, plus per-stream output sizes and chunk counts -- and, with the same inputs resident in HBM,
through ohp_run_streams_device (one walk per stream into bounded regions).  The oracle is the checker here, as in
tests/; nothing under oracle/ is on the product path.
    python profiles/parity_fuzz.py [seconds_budget] [first_seed] > gpurun_out/parity_fuzz.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ohpipeline_b200 import abi, capi, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402
from util import covered_mask  # noqa: E402


def main():
    import torch
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
    first_seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    port = pyoracle.Port()
    ctx = capi.Context(0)
    gens = [("mixed", lambda s: workloads.mixed(n_streams=96, seed=s, max_frames=5000)),
            ("steady_edges", lambda s: workloads.steady_edges(s, n_streams=96)),
            ("config4", lambda s: workloads.config4(n_streams=64, seconds=0.1, seed=s)),
            ("elements", lambda s: workloads.elements(s, n_streams=48, seconds=0.4, illegal=(s % 4 == 0)))]
    tot = {g: {"workloads": 0, "streams": 0, "chunks": 0, "bytes_checked": 0, "asserting_streams_dropped": 0} for g, _ in gens}
    failures = []
    t0 = time.time()
    seed = first_seed
    while time.time() - t0 < budget and not failures:
        for name, make in gens:
            w = make(seed)
            # streams the reference ASSERTs on are found by the host model one at a time and dropped
            keep = []
            for k in range(len(w.streams)):
                s = w.streams[k:k + 1].copy()
                ev = w.events[int(s[0]["first_event"]):int(s[0]["first_event"]) + int(s[0]["num_events"])]
                s[0]["first_event"] = 0
                try:
                    capi.schedule_build(s, ev)
                    keep.append(k)
                except capi.OhpError:
                    tot[name]["asserting_streams_dropped"] += 1
            streams = w.streams[keep].copy()
            inp = port.fill_pcm(w.in_bytes, w.seed)
            rc, want, chunks, _ = port.run(streams, w.events, inp, w.out_bytes)
            assert rc == 0, (name, seed, rc)
            got = np.zeros(w.out_bytes, dtype=np.uint8)
            outb, total = ctx.run_streams_host(streams, w.events, inp, got)
            mask = covered_mask(chunks, w.out_bytes)
            ok = total == len(chunks) and np.array_equal(got[mask], want[mask])
            if not ok:
                failures.append({"generator": name, "seed": seed, "chunks_gpu": int(total), "chunks_oracle": int(len(chunks))})
            # the same batch resident in HBM
            d_s = torch.from_numpy(streams.view(np.uint8).copy()).cuda()
            d_e = (torch.from_numpy(w.events.view(np.uint8).copy()).cuda() if len(w.events) else torch.zeros(32, dtype=torch.uint8, device="cuda"))
            d_in = torch.from_numpy(inp).cuda()
            d_out = torch.zeros(w.out_bytes + 16, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            total_d = ctx.run_streams_device(d_s.data_ptr(), len(streams), d_e.data_ptr(), len(w.events), d_in.data_ptr(), w.in_bytes,
                                             d_out.data_ptr(), w.out_bytes)
            ctx.sync()
            got_d = d_out.cpu().numpy()[:w.out_bytes]
            if not (total_d == len(chunks) and np.array_equal(got_d[mask], want[mask])):
                failures.append({"generator": name, "seed": seed, "path": "ohp_run_streams_device", "chunks_gpu": int(total_d),
                                 "chunks_oracle": int(len(chunks))})
            t = tot[name]
            t["workloads"] += 1; t["streams"] += len(streams); t["chunks"] += int(total); t["bytes_checked"] += int(mask.sum())
        seed += 1
    print(json.dumps({"seconds": round(time.time() - t0, 1), "first_seed": first_seed, "seeds": seed - first_seed, "failures": failures,
                      "paths": ["ohp_run_streams_host", "ohp_run_streams_device"], "totals": tot}))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
