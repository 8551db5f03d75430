#!/usr/bin/env python
"""bench.py -- throughput of the fused ramp + format-convert path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's own CPU code (oracle/_ref)

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1] by default (1024 streams, stereo,
24-bit, 192 kHz, 10 s each, every chunk ramped, packed 24-bit BE out).  At N > 1 every rank processes its own
batch of that size (streams shard with no data-path collective: weak scaling).

Printed on rank 0 as ONE JSON line:
  value      whole-job frames/s ("samples" in the reference's vocabulary = frames) with inputs resident in HBM
  e2e        the same metric through ohp_process_host: pinned HOST buffers, H2D + kernel + D2H inside the timed region
  roofline   algorithmic bytes per launch / average launch duration (CUDA events on the launching stream) vs the
             measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU path (or the C oracle port) on this box's host cores, bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ohpipeline_b200 import abi, workloads  # noqa: E402

METRIC = "PCM samples/sec (fused ramp + format convert)"
UNIT = "samples/s"
DTYPE = "int32 Q15 fixed-point on 16-bit samples (u8 PCM bytes in/out)"


def build_workload(name, streams, seconds):
    if name == "config2":
        return workloads.config2(n_streams=streams or 1024, seconds=seconds or 10.0)
    if name == "config5":
        return workloads.config5(n_streams=streams or 65536, seconds=seconds or 1.0)
    if name == "config3":
        return workloads.config3(n_streams=streams or 4096, seconds=seconds or 1.0,
                                 n_events=int(os.environ.get("OHP_C3_EVENTS", "8")))  # experiment knob: 0 = no ramps, uniform chunks
    if name == "config4":
        return workloads.config4(n_streams=streams or 16384, seconds=seconds or 0.25)
    if name == "mixed":  # the stress version of config4: tiny messages, one-sample caps, one-frame driver blocks
        return workloads.mixed(n_streams=streams or 16384, seed=4, max_frames=int((seconds or 1.0) * 48000))
    if name == "config1":
        return workloads.config1(seconds or 10.0)
    raise SystemExit("unknown workload %r" % name)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload_name, algo_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture of a slice of
    the same workload (profiles/traffic.json), scaled to this launch by algorithmic bytes.  None when no capture exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            rec = json.load(open(p)).get(workload_name)
            if rec:
                return rec["capture_dram_bytes"] / rec["capture_algorithmic_bytes"] * algo_bytes
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []
        self._halt = threading.Event()
        self._active = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if mask & bit and name != "gpu_idle":
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.02)

    def active(self, on):
        (self._active.set if on else self._active.clear)()

    def finish(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": (round(max(self.power), 1) if self.power else None)}


def cpu_reference_run(w, n_streams, threads, steps, warmup):
    """Time the reference's own CPU path (oracle/_ref when present, else the C oracle port) on the first
    n_streams streams of workload w.  Returns (frames/s, ms per step, kind, cores, sample text)."""
    from oracle import pyoracle
    sub = w.streams[:n_streams].copy()
    src_lo = int(sub["src_base"][0])
    src_hi = int(sub["src_base"][-1]) + int(sub["total_frames"][-1]) * int(sub["channels"][-1]) * int(sub["bit_depth"][-1]) // 8
    dst_lo = int(sub["dst_base"][0])
    sub["src_base"] -= src_lo
    sub["dst_base"] -= dst_lo
    in_bytes = src_hi - src_lo
    out_bytes = in_bytes + 4096
    rng = np.random.default_rng(12345)
    inp = rng.integers(0, 256, size=in_bytes, dtype=np.uint8)
    frames = int(sub["total_frames"].sum())
    times = []
    if pyoracle.Ref.available():
        ref = pyoracle.Ref()
        kind = "reference"
        cores = min(threads, n_streams)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            rc, out, _, _ = ref.run(sub, w.events, inp, out_bytes, threads=cores, want_descs=False, want_audio=True)
            dt = time.perf_counter() - t0
            assert rc == 0, rc
            if i >= warmup:
                times.append(dt)
    else:
        port = pyoracle.Port()
        kind = "port"
        cores = 1
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            rc, out, _, _ = port.run(sub, w.events, inp, out_bytes)
            dt = time.perf_counter() - t0
            assert rc == 0, rc
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    sample = "first %d of %d streams of the workload, full length (%d frames/step), %s" % (
        n_streams, len(w.streams), frames,
        "MsgFactory->SetRamp->CreatePlayable->Read(ProcessorPcmBufTest), one MsgFactory per thread" if kind == "reference"
        else "C oracle port, scalar")
    return frames * len(times) / total, 1e3 * total / len(times), kind, cores, sample


def pick_cpu_sample(w, threads, target_s=6.0):
    """Calibrate on one stream-slice, then size the sample so a step costs about target_s seconds of wall time."""
    from oracle import pyoracle
    per_stream_frames = int(w.streams["total_frames"][0])
    if not pyoracle.Ref.available():
        threads = 1
    probe = min(len(w.streams), max(1, threads))
    fps, _, _, _, _ = cpu_reference_run(w, probe, threads, steps=1, warmup=0)
    want_frames = fps * target_s
    n = int(max(probe, min(len(w.streams), want_frames // max(per_stream_frames, 1))))
    return max(1, n)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries ONE JSON line and nothing else: libraries that print there (NCCL's version banner comes from C
    code, at the first collective) are pointed at stderr, the line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=0.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 5)")
    ap.add_argument("--e2e-api", default="run_streams_host", choices=["run_streams_host", "process_host"])
    ap.add_argument("--chunk-frames", type=int, default=0, help="experiment: override the workload's frames per message")
    ap.add_argument("--pad-mb", type=float, default=0.0, help="experiment: spacer allocated between the input and output arenas")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3

    w = build_workload(args.workload, args.streams, args.seconds)
    if args.chunk_frames:
        w.streams["chunk_frames"] = args.chunk_frames
    frames_per_step = w.total_frames
    subsamples_per_step = w.total_subsamples
    cfg = {"workload": w.name, "streams_per_gpu": int(len(w.streams)), "frames_per_step_per_gpu": frames_per_step,
           "l2": "inputs (%.2f GB per step per GPU) are larger than the 126 MB L2" % (w.in_bytes / 1e9),
           "sharding": "independent streams per rank, no data-path collective"}

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # the reference's own CPU implementation; rank 0 alone runs it
        if rank != 0:
            return 0
        import multiprocessing
        threads = multiprocessing.cpu_count()
        n = pick_cpu_sample(w, threads)
        fps, ms, kind, cores, sample = cpu_reference_run(w, n, threads, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------------------
    import torch
    from ohpipeline_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # the image's default prints a version banner on stdout, next to the JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = capi.Context(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()

    # host side: ramp events -> chunk descriptors (product code, threaded over streams)
    t0 = time.perf_counter()
    sched = capi.schedule_build(w.streams, w.events)
    t_sched = time.perf_counter() - t0
    chunks = sched.chunks
    n_chunks = len(chunks)
    payload = int(chunks["bytes"].sum())
    silence = (chunks["flags"] & abi.F_SILENCE) != 0
    in_payload = int(chunks["bytes"][~silence].sum())
    algo_bytes = in_payload + payload + n_chunks * abi.CHUNK_DESC.itemsize

    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + rank)
    d_in = torch.randint(0, 256, (w.in_bytes,), dtype=torch.uint8, device="cuda", generator=g)
    d_pad = torch.zeros(int(args.pad_mb * (1 << 20)) + 1, dtype=torch.uint8, device="cuda")  # noqa: F841 (address spacer)
    d_out = torch.zeros(w.out_bytes, dtype=torch.uint8, device="cuda")
    # descriptors are built ON THE GPU from the stream specs and ramp events (ohp_schedule_{count,emit}_device) and
    # checked against the host model's; the hot path below consumes the device-built array
    d_specs = torch.from_numpy(w.streams.view(np.uint8).copy()).cuda()
    d_events = torch.from_numpy(w.events.view(np.uint8).copy()).cuda() if len(w.events) else torch.zeros(32, dtype=torch.uint8, device="cuda")
    d_begin = torch.zeros(len(w.streams) + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    t_dev = []
    for _ in range(3):
        t0 = time.perf_counter()
        total = ctx.schedule_count_device(d_specs.data_ptr(), len(w.streams), d_events.data_ptr(), len(w.events), d_begin.data_ptr())
        if _ == 0:
            d_desc = torch.empty(max(total, 1) * abi.CHUNK_DESC.itemsize, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
        ctx.schedule_emit_device(d_specs.data_ptr(), len(w.streams), d_events.data_ptr(), len(w.events), d_begin.data_ptr(),
                                 d_desc.data_ptr())
        ctx.sync()
        t_dev.append(time.perf_counter() - t0)
    assert total == n_chunks, (total, n_chunks)
    if not np.array_equal(d_desc.cpu().numpy()[: n_chunks * abi.CHUNK_DESC.itemsize].view(abi.CHUNK_DESC), chunks):
        raise SystemExit("bench.py: descriptors built on the GPU differ from the host model's")
    # an explicit (non-default) stream: a NULL stream argument would select the context's own stream, and the
    # CUDA events below must sit on the stream the kernels are launched on
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0

    def step():
        ctx.process_device(d_desc.data_ptr(), n_chunks, d_in.data_ptr(), w.in_bytes, d_out.data_ptr(), w.out_bytes, st)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.active(True)
    for _ in range(args.warmup):
        step()
    ctx.sync(st)
    launches0 = ctx.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ctx.sync(st)
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    inflight_cap = ctx.inflight_cap()
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = frames_per_step * world / (ms_per_step * 1e-3)

    # result check (cheap, outside the timed region): per-stream checksums exist and are non-trivial
    offs = np.concatenate([w.streams["dst_base"], [w.streams["dst_base"][-1] + sched.stream_out_bytes[-1]]]).astype(np.uint64)
    # streams are padded to 16 bytes: checksum each stream's own bytes only
    sums_t = torch.zeros(len(w.streams), dtype=torch.int64, device="cuda")
    ends = (w.streams["dst_base"] + sched.stream_out_bytes).astype(np.uint64)
    contiguous = bool((ends[:-1] == w.streams["dst_base"][1:]).all())
    if contiguous:
        d_off = torch.from_numpy(offs.view(np.int64).copy()).cuda()
        ctx.checksums_device(d_out.data_ptr(), d_off.data_ptr(), len(w.streams), sums_t.data_ptr(), st)
        ctx.sync(st)
        checksum_of_checksums = int(np.bitwise_xor.reduce(sums_t.cpu().numpy().view(np.uint64)))
    else:
        checksum_of_checksums = None

    # ------------------------------------------------------------------------------------------------------
    # end to end through the C ABI with HOST buffers: H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 5)
        # every rank pins its batch (in + out) in host memory.  Where the box cannot hold all ranks' whole batches the e2e
        # leg runs on the first k streams of each rank's batch (same call, same per-stream work) and says so.
        ns = len(w.streams)
        k = ns
        frac = float(os.environ.get("OHP_E2E_MEM_FRACTION", "0.7"))
        try:
            import psutil
            avail = psutil.virtual_memory().available
            need = (w.in_bytes + w.out_bytes) * world
            if need > frac * avail:
                k = max(1, min(ns, int(frac * avail / need * ns)))
        except ImportError:
            pass
        if dist is not None:
            kt = torch.tensor([k], dtype=torch.int64, device="cuda")
            dist.all_reduce(kt, op=dist.ReduceOp.MIN)
            k = int(kt.item())
        e_streams = w.streams[:k]
        e_in = int(w.streams["src_base"][k]) if k < ns else w.in_bytes
        e_out = int(w.streams["dst_base"][k]) if k < ns else w.out_bytes
        e_chunks = chunks[: int(sched.stream_chunk_begin[k])]
        e_frames = int(e_streams["total_frames"].sum())
        e_payload = int(sched.stream_out_bytes[:k].sum())
        h_in, h_in_ptr = ctx.host_alloc(e_in)
        h_out, h_out_ptr = ctx.host_alloc(e_out)
        ctx.memcpy_d2h(h_in, d_in.data_ptr(), st)
        ctx.sync(st)
        del d_in, d_out, d_desc
        torch.cuda.empty_cache()
        # the call a user makes: stream specs + ramp events + host PCM in, bytes out (descriptors are built on the GPU
        # inside the call, every step); --e2e-api process_host times the descriptor-level entry point instead
        if args.e2e_api == "run_streams_host":
            def e2e_step():
                return ctx.run_streams_host(e_streams, w.events, h_in, h_out)
        else:
            def e2e_step():
                return ctx.process_host(e_chunks, h_in, h_out)
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": e_frames * world * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(e_in + (e_streams.nbytes + w.events.nbytes if args.e2e_api == "run_streams_host"
                                                 else len(e_chunks) * abi.CHUNK_DESC.itemsize)),
               "d2h_bytes_per_step": int(e_payload + (k * 16 + 8 if args.e2e_api == "run_streams_host" else 0)),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "api": ("ohp_run_streams_host (pinned host buffers; specs + events + PCM H2D, descriptors built on the GPU, kernel, "
                       "PCM D2H, sliced by stream and pipelined)" if args.e2e_api == "run_streams_host" else
                       "ohp_process_host (pinned host buffers; descriptors + PCM H2D, kernel, PCM D2H, sliced and pipelined)"),
               "timer": "host wall clock around the synchronous calls, max over ranks"}
        if k < ns:
            e2e["sample"] = "first %d of %d streams per rank (host memory: %.0f GB available for %d ranks)" % (k, ns, avail / 1e9, world)
        e2e_sum = int(h_out[: int(sched.stream_out_bytes[0])].astype(np.uint64).sum())
        e2e["first_stream_byte_sum"] = e2e_sum
        ctx.host_free(h_in_ptr)
        ctx.host_free(h_out_ptr)
    sampler.active(False)
    clocks = sampler.finish()

    # ------------------------------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(args.workload, algo_bytes), "peak_source": peak_src,
                "kernel": "ohp::ramp_convert_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                "launch_ms": ms_per_step}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing
        threads = multiprocessing.cpu_count()
        n = pick_cpu_sample(w, threads, target_s=5.0)
        fps, ms, kind, cores, sample = cpu_reference_run(w, n, threads, steps=3, warmup=1)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "ms_per_step": ms}

    if rank == 0:
        cfg.update({"chunks_per_step_per_gpu": n_chunks, "host_schedule_build_s": round(t_sched, 3),
                    "device_schedule_build_s": round(min(t_dev), 4),
                    "inflight_chunks_per_cta": inflight_cap,
                    "descriptors": "built on the GPU (ohp_schedule_count/emit_device), identical to the host model's"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "subsamples_per_s": subsamples_per_step * world / (ms_per_step * 1e-3),
                "payload_gb_per_s_per_gpu": (in_payload + payload) / (ms_per_step * 1e-3) / 1e9,
                "checksum_of_checksums": checksum_of_checksums}
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
