#!/usr/bin/env python
"""Flywheel ramp generator: device time for a batch of starving streams (CUDA events on the launching stream) next to
the reference's own RampGenerator on one host core.  Usage: python profiles/flywheel_bench.py [n_jobs] [rate] [ch] [bits]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from ohpipeline_b200 import abi, capi  # noqa: E402
from flywheel_util import training_block  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    rate = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
    ch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    bits = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    one = capi.flywheel_job(rate, ch, bits)
    tb = int(one["train_frames"][0]) * 4 * ch
    ob = int(one["out_frames"][0]) * ch * bits // 8
    jobs = np.repeat(one, n)
    jobs["src_off"] = np.arange(n, dtype=np.uint64) * tb
    jobs["dst_off"] = np.arange(n, dtype=np.uint64) * ob
    blocks = np.stack([training_block(rate, ch, "tone", s) for s in range(64)])
    inp = blocks[np.arange(n) % 64].reshape(-1)
    ctx = capi.Context(0)
    d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).cuda()
    d_in = torch.from_numpy(inp).cuda()
    d_out = torch.zeros(n * ob, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    st = stream.cuda_stream
    for _ in range(3):
        ctx.flywheel_device(d_jobs.data_ptr(), n, d_in.data_ptr(), inp.size, d_out.data_ptr(), n * ob, st)
    ctx.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record(stream)
    for _ in range(steps):
        ctx.flywheel_device(d_jobs.data_ptr(), n, d_in.data_ptr(), inp.size, d_out.data_ptr(), n * ob, st)
    e1.record(stream)
    ctx.sync(st)
    ms = e0.elapsed_time(e1) / steps
    line = {"kernel": "ohp::fly::flywheel_kernel", "jobs": n, "rate": rate, "channels": ch, "bit_depth": bits,
            "ms_per_launch": ms, "jobs_per_s": n / (ms * 1e-3), "generated_frames_per_s": n * int(one["out_frames"][0]) / (ms * 1e-3),
            "algorithmic_gb_per_s": n * (tb + ob) / (ms * 1e-3) / 1e9}
    try:
        from oracle import pyoracle
        if pyoracle.Ref.available():
            ref = pyoracle.Ref()
            t0 = time.perf_counter()
            k = 200
            for i in range(k):
                ref.flywheel(rate, ch, bits, abi.RAMP_MAX, blocks[i % 64])
            dt = time.perf_counter() - t0
            line["reference_jobs_per_s_one_core"] = k / dt
            line["reference_note"] = "real RampGenerator incl. thread hand-off and the ramped Read of every block"
    except Exception as e:  # noqa: BLE001
        line["reference_error"] = str(e)
    import json
    print(json.dumps(line))


if __name__ == "__main__":
    main()
