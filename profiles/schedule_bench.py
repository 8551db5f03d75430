"""Device-side schedule builder (ohp_schedule_count_device / ohp_schedule_emit_device): time per pass, per workload,
next to the host model on all host cores; descriptors checked against the host model's.
Run on the GPU box:  python profiles/schedule_bench.py > gpurun_out/schedule_bench.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ohpipeline_b200 import abi, capi, workloads  # noqa: E402


def main():
    ctx = capi.Context(0)
    cases = [
        ("config2_1024x10s", lambda: workloads.config2()),
        ("config3_4096x1s", lambda: workloads.config3()),
        ("config5_65536x0.25s", lambda: workloads.config5(seconds=0.25)),
        ("mixed_16384", lambda: workloads.mixed(n_streams=16384, seed=4, max_frames=48000)),
    ]
    only = sys.argv[1:]
    out = {}
    for name, make in cases:
        if only and name not in only:
            continue
        w = make()
        t0 = time.perf_counter()
        host = capi.schedule_build(w.streams, w.events)
        t_host = time.perf_counter() - t0
        d_specs = torch.from_numpy(w.streams.view(np.uint8).copy()).cuda()
        d_events = (torch.from_numpy(w.events.view(np.uint8).copy()).cuda() if len(w.events)
                    else torch.zeros(32, dtype=torch.uint8, device="cuda"))
        d_begin = torch.zeros(len(w.streams) + 1, dtype=torch.int64, device="cuda")
        d_desc = torch.empty(max(len(host.chunks), 1) * abi.CHUNK_DESC.itemsize, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        tc, te = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            total = ctx.schedule_count_device(d_specs.data_ptr(), len(w.streams), d_events.data_ptr(), len(w.events),
                                              d_begin.data_ptr())
            t1 = time.perf_counter()
            ctx.schedule_emit_device(d_specs.data_ptr(), len(w.streams), d_events.data_ptr(), len(w.events),
                                     d_begin.data_ptr(), d_desc.data_ptr())
            ctx.sync()
            t2 = time.perf_counter()
            tc.append(t1 - t0)
            te.append(t2 - t1)
        same = total == len(host.chunks) and np.array_equal(
            d_desc.cpu().numpy()[: total * abi.CHUNK_DESC.itemsize].view(abi.CHUNK_DESC), host.chunks)
        out[name] = {"streams": len(w.streams), "chunks": int(total), "count_ms": round(min(tc) * 1e3, 3),
                     "emit_ms": round(min(te) * 1e3, 3), "host_model_ms": round(t_host * 1e3, 1),
                     "host_threads": os.cpu_count(), "identical_to_host_model": bool(same)}
        del d_specs, d_events, d_begin, d_desc
    print(json.dumps(out))


if __name__ == "__main__":
    main()
