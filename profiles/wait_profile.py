#!/usr/bin/env python
"""Where does each warp role of ramp_convert_kernel spend its time?  Needs a library built with -DOHP_PROFILE_WAITS:

    nvcc ... -DOHP_PROFILE_WAITS -o build/libohp_prof.so ohpipeline_b200/csrc/ohp_capi.cu
    OHP_LIB_CUDA=$PWD/build/libohp_prof.so python profiles/wait_profile.py [config2|config5|config3] [streams] [seconds]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ohpipeline_b200 import capi, workloads  # noqa: E402
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 512
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
w = bench.build_workload(name, streams, seconds)
sched = capi.schedule_build(w.streams, w.events)
ctx = capi.Context(0)
d_in = torch.randint(0, 256, (w.in_bytes,), dtype=torch.uint8, device="cuda")
d_out = torch.zeros(w.out_bytes, dtype=torch.uint8, device="cuda")
d_desc = torch.from_numpy(sched.chunks.view(np.uint8).copy()).cuda()
torch.cuda.synchronize()
stream = torch.cuda.Stream()
st = stream.cuda_stream
L = capi.cuda_lib()
L.ohp_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
buf = (C.c_uint32 * 16)()
for _ in range(3):
    ctx.process_device(d_desc.data_ptr(), len(sched.chunks), d_in.data_ptr(), w.in_bytes, d_out.data_ptr(), w.out_bytes, st)
ctx.sync(st)
L.ohp_debug_counters(ctx._h, buf)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
ctx.process_device(d_desc.data_ptr(), len(sched.chunks), d_in.data_ptr(), w.in_bytes, d_out.data_ptr(), w.out_bytes, st)
ev1.record(stream)
ctx.sync(st)
L.ohp_debug_counters(ctx._h, buf)
v = [buf[i] * 4096.0 for i in range(16)]
life = v[8]
print("%s: kernel %.3f ms, %d chunks; CTA-lifetime cycles summed over CTAs: %.3g" % (w.name, ev0.elapsed_time(ev1), len(sched.chunks), life))
for label, i in (("loader blocked: ring full, waiting for a slot to be released", 4),
                 ("consumer warp 0 blocked: waiting for its chunk's data", 5),
                 ("consumer warp 0 blocked: waiting for its bulk store to finish reading the slot", 6),
                 ("consumer warp 0: transform (or shift) of the chunk", 9),
                 ("consumer warp 0: fence.proxy.async + __syncwarp", 10),
                 ("consumer warp 0: issuing the store (ragged bytes + TMA)", 11)):
    print("  %-82s %5.1f%% of CTA lifetime" % (label, 100.0 * v[i] / max(life, 1)))
rounds, infl = buf[7], buf[14] * 16.0
if rounds:
    print("  loader: %.2f chunks placed per issue round, %.1f chunks in flight when a round starts (cap %d)"
          % (len(sched.chunks) / rounds, infl / rounds, ctx.inflight_cap()))
