"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
  port : oracle/libohp_oracle.so   -- plain-C restatement (ohp_oracle.c)
  ref  : oracle/_ref/libohref.so   -- the reference's own Msg.cpp compiled unmodified + ref_harness.cpp
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

import sys
if REPO not in sys.path:
    sys.path.insert(0, REPO)
from ohpipeline_b200 import abi  # plain dtypes/constants only


class ScheduleResult(C.Structure):
    _fields_ = [("chunks", C.c_void_p), ("info", C.c_void_p), ("num_chunks", C.c_size_t),
                ("stream_chunk_begin", C.POINTER(C.c_uint64)), ("stream_out_bytes", C.POINTER(C.c_uint64))]


def build(force=False):
    """Build the port (always) and the linked reference (when /root/reference is mounted)."""
    port = os.path.join(HERE, "libohp_oracle.so")
    if force or not os.path.exists(port) or not os.path.exists(os.path.join(HERE, "_ref", "libohref.so")):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _Lib:
    """Common surface of the two oracle libraries (function prefix differs)."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        f = getattr(L, prefix + "jiffies_per_sample"); f.restype = C.c_uint32; f.argtypes = [C.c_uint32]
        f = getattr(L, prefix + "ramp_set"); f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        f = getattr(L, prefix + "ramp_split"); f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        f = getattr(L, prefix + "median_multiplier"); f.restype = C.c_uint32; f.argtypes = [C.c_void_p]
        f = getattr(L, prefix + "process_chunks"); f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]

    def jiffies_per_sample(self, rate):
        return getattr(self.lib, self.prefix + "jiffies_per_sample")(rate)

    def ramp_set(self, ramp, start, frag, dur, direction):
        """ramp = (start,end,direction,enabled).  Returns (rc, ramp, split, split_pos)."""
        r = np.array([tuple(ramp)], dtype=abi.RAMP)
        s = np.zeros(1, dtype=abi.RAMP)
        pos = C.c_uint32(0)
        rc = getattr(self.lib, self.prefix + "ramp_set")(_ptr(r), start, frag, dur, direction, _ptr(s), C.byref(pos))
        return rc, tuple(int(x) for x in r[0]), tuple(int(x) for x in s[0]), pos.value

    def ramp_split(self, ramp, new_size, cur_size):
        r = np.array([tuple(ramp)], dtype=abi.RAMP)
        rem = np.zeros(1, dtype=abi.RAMP)
        rc = getattr(self.lib, self.prefix + "ramp_split")(_ptr(r), new_size, cur_size, _ptr(rem))
        return rc, tuple(int(x) for x in r[0]), tuple(int(x) for x in rem[0])

    def median_multiplier(self, ramp):
        r = np.array([tuple(ramp)], dtype=abi.RAMP)
        return getattr(self.lib, self.prefix + "median_multiplier")(_ptr(r))

    def process_chunks(self, descs, inp, out_bytes):
        """Apply descriptors to `inp` (uint8 array); returns (rc, out uint8 array)."""
        descs = np.ascontiguousarray(descs, dtype=abi.CHUNK_DESC)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.zeros(int(out_bytes), dtype=np.uint8)
        rc = getattr(self.lib, self.prefix + "process_chunks")(
            _ptr(descs), len(descs), _ptr(inp), inp.size, _ptr(out), out.size)
        return int(rc), out

    # ---- flywheel ramp generator (both libraries export burgs_method / feedback_model) ----
    def burgs_method(self, samples, degree):
        """FlywheelRamper::BurgsMethod on int16 samples; returns the `degree` coefficients."""
        samples = np.ascontiguousarray(samples, dtype=np.int16)
        out = np.zeros(degree, dtype=np.int16)
        h = np.zeros(degree, dtype=np.int16)
        per = np.zeros(len(samples), dtype=np.int16)
        pef = np.zeros(len(samples), dtype=np.int16)
        f = getattr(self.lib, self.prefix + "burgs_method")
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f(_ptr(samples), len(samples), degree, _ptr(out), _ptr(h), _ptr(per), _ptr(pef))
        return out

    def _collect(self, res, n_streams, free):
        n = res.num_chunks
        chunks = np.zeros(n, dtype=abi.CHUNK_DESC)
        info = np.zeros(n, dtype=abi.CHUNK_INFO)
        if n:
            C.memmove(_ptr(chunks), res.chunks, n * abi.CHUNK_DESC.itemsize)
            C.memmove(_ptr(info), res.info, n * abi.CHUNK_INFO.itemsize)
        begin = np.array([res.stream_chunk_begin[i] for i in range(n_streams + 1)], dtype=np.uint64)
        outb = np.array([res.stream_out_bytes[i] for i in range(n_streams)], dtype=np.uint64)
        free(C.byref(res))
        return chunks, info, begin, outb


class Port(_Lib):
    def __init__(self):
        path = os.path.join(HERE, "libohp_oracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path, "ohpo_")
        L = self.lib
        L.ohpo_schedule_run.restype = C.c_int
        L.ohpo_schedule_run.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(ScheduleResult)]
        L.ohpo_schedule_result_free.argtypes = [C.POINTER(ScheduleResult)]
        L.ohpo_checksum.restype = C.c_uint64
        L.ohpo_checksum.argtypes = [C.c_void_p, C.c_uint64]
        L.ohpo_fill_pcm.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.ohpo_unpack_to_be.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.ohpo_chunk_out_bytes.restype = C.c_uint32
        L.ohpo_chunk_out_bytes.argtypes = [C.c_void_p]
        self.ramp_array = np.ctypeslib.as_array((C.c_uint32 * 512).in_dll(L, "ohpo_ramp_array")).copy()

    def feedback_model(self, descale, coeff_format, data_format, out_format, coeffs, samples, count):
        """FeedbackModel: `count` outputs from the coefficients and initial states (int32 arrays)."""
        coeffs = np.array(coeffs, dtype=np.int64).astype(np.uint32).view(np.int32).copy()
        samples = np.array(samples, dtype=np.int64).astype(np.uint32).view(np.int32).copy()

        class FB(C.Structure):
            _fields_ = [("coeffs", C.c_void_p), ("samples", C.c_void_p), ("state_count", C.c_uint32),
                        ("descale_bits", C.c_uint32), ("coeff_format", C.c_uint32), ("shift", C.c_int32)]
        fb = FB()
        L = self.lib
        L.ohpo_feedback_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ohpo_feedback_next.restype = C.c_int32
        L.ohpo_feedback_next.argtypes = [C.c_void_p]
        L.ohpo_feedback_init(C.byref(fb), len(coeffs), descale, coeff_format, data_format, out_format, _ptr(coeffs), _ptr(samples))
        return np.array([L.ohpo_feedback_next(C.byref(fb)) for _ in range(count)], dtype=np.int32)

    def flywheel(self, jobs, inp, out_bytes):
        """FlywheelRamperManager::Ramp + RampGenerator::ProcessFragment per job; returns (rc, out)."""
        jobs = np.ascontiguousarray(jobs, dtype=abi.FLYWHEEL_JOB)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.zeros(int(out_bytes), dtype=np.uint8)
        L = self.lib
        L.ohpo_flywheel.restype = C.c_int64
        L.ohpo_flywheel.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        rc = L.ohpo_flywheel(_ptr(jobs), len(jobs), _ptr(inp), inp.size, _ptr(out), out.size)
        return int(rc), out

    def flywheel_ramp_chunks(self, job, current_ramp, src_off, dst_off):
        """RampGenerator::Start/EndBlock: (descs, final_ramp) or (None, None) where the reference would ASSERT."""
        job = np.ascontiguousarray(job, dtype=abi.FLYWHEEL_JOB).reshape(1)
        descs = np.zeros(64, dtype=abi.CHUNK_DESC)
        final = C.c_uint32(0)
        L = self.lib
        L.ohpo_flywheel_ramp_chunks.restype = C.c_int
        L.ohpo_flywheel_ramp_chunks.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        n = L.ohpo_flywheel_ramp_chunks(_ptr(job), current_ramp, src_off, dst_off, _ptr(descs), len(descs), C.byref(final))
        if n < 0:
            return None, None
        return descs[:n].copy(), int(final.value)

    def schedule_run(self, streams, events):
        """Message-model restatement: returns (rc, chunks, info, stream_chunk_begin, stream_out_bytes)."""
        streams = np.ascontiguousarray(streams, dtype=abi.STREAM_SPEC)
        events = np.ascontiguousarray(events, dtype=abi.RAMP_EVENT)
        res = ScheduleResult()
        rc = self.lib.ohpo_schedule_run(_ptr(streams), len(streams), _ptr(events), len(events), C.byref(res))
        if rc != 0:
            return rc, None, None, None, None
        return (0,) + self._collect(res, len(streams), self.lib.ohpo_schedule_result_free)

    def run(self, streams, events, inp, out_bytes):
        """schedule_run + process_chunks: the port's end-to-end output for a batch."""
        rc, chunks, info, begin, outb = self.schedule_run(streams, events)
        if rc != 0:
            return rc, None, None, None
        rc2, out = self.process_chunks(chunks, inp, out_bytes)
        return int(rc2), out, chunks, info

    def checksum(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.lib.ohpo_checksum(_ptr(data), data.size))

    def fill_pcm(self, nbytes, seed):
        buf = np.empty(int(nbytes), dtype=np.uint8)
        self.lib.ohpo_fill_pcm(_ptr(buf), buf.size, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF))
        return buf

    def fill_streams(self, streams, in_bytes, seed_base, first_stream_id=0):
        """The bytes ohp_fill_streams_device generates: stream s = ohpo_fill_pcm seeded seed_base | (first_stream_id + s)
        at its src_base.  Bytes between streams are zero."""
        buf = np.zeros(int(in_bytes), dtype=np.uint8)
        base = buf.ctypes.data
        for s in range(len(streams)):
            sp = streams[s]
            n = int(sp["total_frames"]) * int(sp["channels"]) * (int(sp["bit_depth"]) // 8)
            seed = (int(seed_base) | (int(first_stream_id) + s)) & 0xFFFFFFFFFFFFFFFF
            self.lib.ohpo_fill_pcm(C.c_void_p(base + int(sp["src_base"])), C.c_uint64(n), C.c_uint64(seed))
        return buf

    def unpack_to_be(self, src, bits, little_endian):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        dst = np.empty_like(src)
        self.lib.ohpo_unpack_to_be(_ptr(src), _ptr(dst), src.size, bits, int(little_endian))
        return dst


class Ref(_Lib):
    """The reference's own code.  Present in the build container and wherever the prebuilt .so travelled."""

    PATH = os.path.join(HERE, "_ref", "libohref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        super().__init__(self.PATH, "ref_")
        L = self.lib
        L.ref_schedule_run.restype = C.c_int
        L.ref_schedule_run.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                       C.POINTER(ScheduleResult), C.c_int]
        L.ref_ramp_array.restype = C.POINTER(C.c_uint32)
        L.ref_ramp_array_count.restype = C.c_uint32
        L.ref_hardware_threads.restype = C.c_int
        self._libc_free = C.CDLL(None).free
        self._libc_free.argtypes = [C.c_void_p]
        n = L.ref_ramp_array_count()
        self.ramp_array = np.array([L.ref_ramp_array()[i] for i in range(n)], dtype=np.uint32)

    def feedback_model(self, descale, coeff_format, data_format, out_format, coeffs, samples, count):
        coeffs = np.array(coeffs, dtype=np.int64).astype(np.uint32).view(np.int32).copy()
        samples = np.array(samples, dtype=np.int64).astype(np.uint32).view(np.int32).copy()
        out = np.zeros(count, dtype=np.int32)
        f = self.lib.ref_feedback_model
        f.restype = None
        f.argtypes = [C.c_uint32] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        f(len(coeffs), descale, coeff_format, data_format, out_format, _ptr(coeffs), _ptr(samples), _ptr(out), count)
        return out

    def flywheel(self, rate, channels, bits, current_ramp, training):
        """The real RampGenerator.  Returns (rc, raw, ramped, descs, info, final_ramp)."""
        training = np.ascontiguousarray(training, dtype=np.uint8)
        cap = 20 * 400 * channels * 4 + 64
        raw = np.zeros(cap, dtype=np.uint8)
        ramped = np.zeros(cap, dtype=np.uint8)
        descs = np.zeros(64, dtype=abi.CHUNK_DESC)
        info = np.zeros(64, dtype=abi.CHUNK_INFO)
        nbytes, nd, final = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        f = self.lib.ref_flywheel
        f.restype = C.c_int
        f.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                      C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        rc = f(rate, channels, bits, current_ramp, _ptr(training), training.size, _ptr(raw), _ptr(ramped), cap,
               C.byref(nbytes), _ptr(descs), _ptr(info), len(descs), C.byref(nd), C.byref(final))
        if rc != 0:
            return rc, None, None, None, None, None
        return 0, raw[:nbytes.value].copy(), ramped[:nbytes.value].copy(), descs[:nd.value].copy(), info[:nd.value].copy(), int(final.value)

    def flywheel_input(self, descs, inp, rate, jiffies):
        """The real FlywheelInput::Prepare over messages made from `descs`.  Returns (rc, planar bytes)."""
        descs = np.ascontiguousarray(descs, dtype=abi.CHUNK_DESC)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        planar = np.zeros(16384, dtype=np.uint8)
        n = C.c_uint32(0)
        f = self.lib.ref_flywheel_input
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        rc = f(_ptr(descs), len(descs), _ptr(inp), rate, jiffies, _ptr(planar), planar.size, C.byref(n))
        return rc, planar[:n.value].copy()

    def aggregate(self, rate, channels, bits, frames):
        """The real DecodedAudioAggregator: frame counts of the messages it passes on for the given input messages."""
        frames = np.ascontiguousarray(frames, dtype=np.uint32)
        out = np.zeros(len(frames) + 4, dtype=np.uint32)
        f = self.lib.ref_aggregate
        f.restype = C.c_int
        f.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        n = f(rate, channels, bits, _ptr(frames), len(frames), _ptr(out), len(out))
        return None if n < 0 else out[:n].copy()

    def process_chunks_sinks(self, descs, inp, out_bytes, fill=0):
        """MsgPlayable::Read of every descriptor through the REAL sink its out_fmt names (ProcessorPcmBufTest, the extracted
        ProcessorPcmSwpEndianPacked, the extracted Songcast Sender).  Returns (rc, out, bytes each chunk's sink was left holding)."""
        descs = np.ascontiguousarray(descs, dtype=abi.CHUNK_DESC)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.full(int(out_bytes), fill, dtype=np.uint8)
        sizes = np.zeros(len(descs), dtype=np.uint32)
        f = self.lib.ref_process_chunks_sinks
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        rc = f(_ptr(descs), len(descs), _ptr(inp), inp.size, _ptr(out), out.size, _ptr(sizes))
        return int(rc), out, sizes

    def container_decode(self, data, max_bit_depth=32):
        """The reference's own CodecWav / CodecAifc / CodecAiff on the container bytes, behind a fake CodecController.
        Returns (ohp_container_status, CONTAINER_INFO record, OutputAudioPcm byte counts (AIFF / AIFC only))."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        info = np.zeros(1, dtype=abi.CONTAINER_INFO)
        cap = max(16, len(buf) // 64 + 16)
        reads = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        f = self.lib.ref_container_decode
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        rc = f(_ptr(buf) if buf.size else None, buf.size, max_bit_depth, _ptr(info), _ptr(reads), cap, C.byref(n))
        return int(rc), info[0], reads[:min(n.value, cap)].copy()

    def elements_run(self, streams, events, inp, out_bytes, want_audio=True):
        """Every stream through the reference's own element OBJECTS (Ramper, Muter, StarvationRamper; oracle/ref_elements.cpp),
        driven at the positions the element-level OHP_EV_* events name.  Returns (rc, out, chunks, info, chunk_begin,
        stream_out_bytes, flywheel messages played per stream); rc -1: the reference ASSERTs, -2: not stageable."""
        streams = np.ascontiguousarray(streams, dtype=abi.STREAM_SPEC)
        events = np.ascontiguousarray(events, dtype=abi.RAMP_EVENT)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.zeros(int(out_bytes), dtype=np.uint8) if want_audio else None
        gen = np.zeros(max(len(streams), 1), dtype=np.uint32)
        res = ScheduleResult()
        f = self.lib.ref_elements_run
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(ScheduleResult), C.c_void_p]
        rc = f(_ptr(streams), len(streams), _ptr(events), len(events), _ptr(inp), _ptr(out) if want_audio else None, C.byref(res), _ptr(gen))
        if rc != 0:
            return rc, None, None, None, None, None, None
        chunks, info, begin, outb = self._collect(res, len(streams), self._free_result)
        return 0, out, chunks, info, begin, outb, gen[:len(streams)].copy()

    def elements_generated_audio(self, stream, events, inp, cap=1 << 22):
        """ONE stream through the reference's element objects; returns (rc, the bytes a driver reads from the flywheel messages
        its StarvationRamper plays when it starves -- every starvation, one after the other --, the element's ramp value at each)."""
        stream = np.ascontiguousarray(stream, dtype=abi.STREAM_SPEC)
        events = np.ascontiguousarray(events, dtype=abi.RAMP_EVENT)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        audio = np.zeros(cap, dtype=np.uint8)
        ramps = np.zeros(64, dtype=np.uint32)
        n = C.c_uint32(0)
        f = self.lib.ref_elements_generated_audio
        f.restype = C.c_long
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
        got = f(_ptr(stream), _ptr(events), _ptr(inp), _ptr(audio), cap, _ptr(ramps), len(ramps), C.byref(n))
        if got < 0:
            return int(got), None, None
        assert got <= cap
        return 0, audio[:got].copy(), ramps[:n.value].copy()

    def volume_ramper(self, enabled, ramps, kinds):
        """The real VolumeRamper element on one message per ramp (kinds: 0 PCM, 1 silence).  Returns (multipliers handed to
        IVolumeRamper::ApplyVolumeMultiplier, the ramp each message leaves with)."""
        r = np.array([tuple(x) for x in ramps], dtype=abi.RAMP)
        k = np.ascontiguousarray(kinds, dtype=np.uint32)
        mult = np.zeros(len(r) + 4, dtype=np.uint32)
        after = np.zeros(len(r), dtype=abi.RAMP)
        f = self.lib.ref_volume_ramper
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        n = f(int(bool(enabled)), _ptr(r), _ptr(k), len(r), _ptr(mult), _ptr(after))
        assert n >= 0
        return mult[:n].copy(), after

    def _free_result(self, res_ref):
        res = res_ref._obj
        self._libc_free(res.chunks)
        self._libc_free(res.info)
        self._libc_free(C.cast(res.stream_chunk_begin, C.c_void_p))
        self._libc_free(C.cast(res.stream_out_bytes, C.c_void_p))

    def hardware_threads(self):
        return int(self.lib.ref_hardware_threads())

    def run(self, streams, events, inp, out_bytes, threads=1, want_descs=True, want_audio=True, want_sizes=False):
        """The real MsgFactory -> SetRamp -> CreatePlayable -> Read(ProcessorPcmBufTest) path.
        Returns (rc, out, chunks, info); with want_sizes (needs want_descs) also each stream's output bytes."""
        streams = np.ascontiguousarray(streams, dtype=abi.STREAM_SPEC)
        events = np.ascontiguousarray(events, dtype=abi.RAMP_EVENT)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.zeros(int(out_bytes), dtype=np.uint8) if want_audio else None
        res = ScheduleResult()
        rc = self.lib.ref_schedule_run(_ptr(streams), len(streams), _ptr(events), len(events), _ptr(inp),
                                       _ptr(out) if want_audio else None,
                                       C.byref(res) if want_descs else None, threads)
        if rc != 0:
            return (rc, None, None, None, None) if want_sizes else (rc, None, None, None)
        chunks = info = outb = None
        if want_descs:
            chunks, info, _, outb = self._collect(res, len(streams), self._free_result)
        return (0, out, chunks, info, outb) if want_sizes else (0, out, chunks, info)
