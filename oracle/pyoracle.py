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

    def _free_result(self, res_ref):
        res = res_ref._obj
        self._libc_free(res.chunks)
        self._libc_free(res.info)
        self._libc_free(C.cast(res.stream_chunk_begin, C.c_void_p))
        self._libc_free(C.cast(res.stream_out_bytes, C.c_void_p))

    def hardware_threads(self):
        return int(self.lib.ref_hardware_threads())

    def run(self, streams, events, inp, out_bytes, threads=1, want_descs=True, want_audio=True):
        """The real MsgFactory -> SetRamp -> CreatePlayable -> Read(ProcessorPcmBufTest) path.
        Returns (rc, out, chunks, info)."""
        streams = np.ascontiguousarray(streams, dtype=abi.STREAM_SPEC)
        events = np.ascontiguousarray(events, dtype=abi.RAMP_EVENT)
        inp = np.ascontiguousarray(inp, dtype=np.uint8)
        out = np.zeros(int(out_bytes), dtype=np.uint8) if want_audio else None
        res = ScheduleResult()
        rc = self.lib.ref_schedule_run(_ptr(streams), len(streams), _ptr(events), len(events), _ptr(inp),
                                       _ptr(out) if want_audio else None,
                                       C.byref(res) if want_descs else None, threads)
        if rc != 0:
            return rc, None, None, None
        chunks = info = None
        if want_descs:
            chunks, info, _, _ = self._collect(res, len(streams), self._free_result)
        return 0, out, chunks, info
