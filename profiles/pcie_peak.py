"""PCIe ceiling on the bench box: what ohp_process_host's end-to-end number can at best reach.
H2D alone, D2H alone, both directions at once (pinned buffers from ohp_host_alloc, i.e. NUMA-local), in 48 MB and in
whole-buffer copies; and the host-side cost of validating/slicing configs[1]'s 2.05 M descriptors.
Run on the GPU box:  python profiles/pcie_peak.py > gpurun_out/pcie_peak.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ohpipeline_b200 import capi  # noqa: E402


def main():
    ctx = capi.Context(0)
    n = 4 << 30
    h_a, p_a = ctx.host_alloc(n)
    h_b, p_b = ctx.host_alloc(n)
    h_a[:] = 1
    h_b[:] = 2
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def run(name, do_in, do_out, piece):
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for off in range(0, n, piece):
                m = min(piece, n - off)
                if do_in:
                    ctx.memcpy_h2d(d_a.data_ptr() + off, h_a[off:off + m], stream=s_in.cuda_stream)
                if do_out:
                    ctx.memcpy_d2h(h_b[off:off + m], d_b.data_ptr() + off, stream=s_out.cuda_stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        moved = n * (int(do_in) + int(do_out))
        res[name] = {"seconds": best, "gb_per_s_total": moved / best / 1e9, "gb_per_s_per_direction": n / best / 1e9}

    for piece, tag in ((n, "whole"), (48 << 20, "48MB")):
        run("h2d_" + tag, True, False, piece)
        run("d2h_" + tag, False, True, piece)
        run("both_" + tag, True, True, piece)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
