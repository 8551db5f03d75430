#!/usr/bin/env python
"""End to end through ohp_multi_run_streams_host (include/ohp_multi.h): ONE process, one host thread + context per device,
pinned host buffers in, pinned host buffers out -- beside the same batch through ohp_run_streams_host on one device.

    python profiles/multi_e2e.py [seconds_of_audio] [repeats] > gpurun_out/r02_multi_e2e.json

Workload: BASELINE configs[1]'s streams (1024 x 2ch/24/192k, every chunk ramped) at `seconds_of_audio` each (default 2:
2.36 GB each way).  Prints one JSON line: per device list the best and median wall-clock time of `repeats` calls, GB/s over
PCIe both directions summed, frames/s, and whether bytes and per-stream checksums equal the one-device call's."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from ohpipeline_b200 import capi, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402  (checker only: seeded PCM and the checksum definition)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    port = pyoracle.Port()
    w = workloads.config2(n_streams=1024, seconds=seconds)
    n_dev = capi.device_count()
    lists = [[0]] + ([list(range(k)) for k in (2, 4, 8) if k <= n_dev] if n_dev > 1 else [[0, 0]])
    out = {"workload": w.name, "in_bytes": w.in_bytes, "out_bytes": w.out_bytes, "frames": w.total_frames, "devices_visible": n_dev, "runs": []}
    one = capi.Context(0)
    h_in, p_in = one.host_alloc(w.in_bytes)
    h_out, p_out = one.host_alloc(w.out_bytes)
    h_in[:] = port.fill_pcm(w.in_bytes, w.seed)
    # the one-device call: the bytes to compare with, and its own time
    times = []
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        want_outb, want_total = one.run_streams_host(w.streams, w.events, h_in, h_out)
        times.append(time.perf_counter() - t0)
    want = h_out.copy()
    sample = list(range(0, len(w.streams), max(1, len(w.streams) // 16)))
    want_sums = {s: port.checksum(want[int(w.streams[s]["dst_base"]):int(w.streams[s]["dst_base"]) + int(want_outb[s])]) for s in sample}
    moved = w.in_bytes + int(want_outb.sum())

    def line(api, devices, ts, same):
        ts = sorted(ts[1:])  # the first call sizes every device's buffers
        return {"api": api, "devices": devices, "best_ms": 1e3 * ts[0], "median_ms": 1e3 * ts[len(ts) // 2],
                "pcie_gbs_best": moved / ts[0] / 1e9, "frames_per_s_best": w.total_frames / ts[0], "same_bytes_and_checksums": same}

    out["runs"].append(line("ohp_run_streams_host", [0], times, True))
    for devices in lists:
        m = capi.MultiContext(devices)
        times, same = [], True
        for _ in range(repeats + 1):
            h_out[:] = 0
            t0 = time.perf_counter()
            outb, sums, total = m.run_streams_host(w.streams, w.events, h_in, h_out)
            times.append(time.perf_counter() - t0)
            same = same and total == want_total and np.array_equal(outb, want_outb) and np.array_equal(h_out, want) \
                and all(int(sums[s]) == want_sums[s] for s in sample)
        out["runs"].append(line("ohp_multi_run_streams_host", devices, times, bool(same)))
        m.close()
    one.host_free(p_in)
    one.host_free(p_out)
    one.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
