"""The C-ABI libraries load without a GPU and export every symbol include/*.h declares; struct layouts match."""
import ctypes
import os
import re

import numpy as np
import pytest

from ohpipeline_b200 import abi, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ohp_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    lib = capi.cuda_lib()
    names = declared_functions("ohp_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_exports_the_device_schedule_builder():
    lib = capi.cuda_lib()
    names = declared_functions("ohp_schedule_device.h")
    assert names == ["ohp_fill_streams_device", "ohp_run_streams_device", "ohp_run_streams_host", "ohp_schedule_count_device",
                     "ohp_schedule_emit_device"]
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_exports_the_flywheel_generator():
    lib = capi.cuda_lib()
    names = declared_functions("ohp_flywheel.h")
    assert names == ["ohp_flywheel_device", "ohp_flywheel_out_bytes", "ohp_flywheel_validate"]
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_exports_the_multi_device_driver_and_it_fails_loudly_without_a_gpu():
    lib = capi.cuda_lib()
    names = declared_functions("ohp_multi.h")
    assert names == ["ohp_multi_context", "ohp_multi_create", "ohp_multi_destroy", "ohp_multi_host_alloc", "ohp_multi_host_free",
                     "ohp_multi_last_error", "ohp_multi_num_devices", "ohp_multi_run_streams_host", "ohp_multi_shard"]
    for n in names:
        assert hasattr(lib, n), n
    # the block rule is sharding.py's (one process per GPU) -- contiguous, sizes differing by at most one, any sizes
    from ohpipeline_b200 import sharding
    for n_streams in (0, 1, 7, 1000, 65536, 65537, 2**40 + 3):
        for n_devices in (1, 2, 3, 8, 64):
            at = 0
            for g in range(n_devices):
                first, count = capi.multi_shard(n_streams, n_devices, g)
                assert (first, first + count) == sharding.shard_range(n_streams, n_devices, g) and first == at
                at += count
            assert at == n_streams
    if capi.device_count() == 0:
        with pytest.raises(capi.OhpError) as e:
            capi.MultiContext([0, 0])
        assert e.value.status in (abi.E_NO_DEVICE, abi.E_CUDA) and "device 0" in str(e.value)
    with pytest.raises(capi.OhpError):
        capi.MultiContext([])


def test_host_library_exports_the_container_front_end():
    lib = capi.host_lib()
    names = declared_functions("ohp_container.h")
    assert names == ["ohp_codec_message_frames", "ohp_container_parse", "ohp_container_stream_spec"]
    for n in names:
        assert hasattr(lib, n), n


def test_host_library_exports_every_declared_symbol():
    lib = capi.host_lib()
    names = declared_functions("ohp_schedule.h")
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version_and_constants():
    assert capi.cuda_lib().ohp_abi_version() == abi.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "ohp_b200.h")).read()
    assert "#define OHP_RAMP_MAX 16384u" in hdr
    assert "#define OHP_MAX_PCM_CHUNK_BYTES 9216u" in hdr


def test_struct_layout_matches_header():
    # compile a tiny C program that prints sizeof/offsetof and compare with the numpy dtypes
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "ohp_schedule.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("ohp_chunk_desc %zu\n", sizeof(ohp_chunk_desc));
  P(ohp_chunk_desc, src_off); P(ohp_chunk_desc, dst_off); P(ohp_chunk_desc, bytes); P(ohp_chunk_desc, ramp_start);
  P(ohp_chunk_desc, ramp_end); P(ohp_chunk_desc, attenuation); P(ohp_chunk_desc, bit_depth); P(ohp_chunk_desc, channels);
  P(ohp_chunk_desc, flags); P(ohp_chunk_desc, out_fmt); P(ohp_chunk_desc, aux);
  printf("ohp_ramp_event %zu\n", sizeof(ohp_ramp_event));
  P(ohp_ramp_event, at_jiffies); P(ohp_ramp_event, stage); P(ohp_ramp_event, op); P(ohp_ramp_event, arg);
  printf("ohp_stream_spec %zu\n", sizeof(ohp_stream_spec));
  P(ohp_stream_spec, sample_rate); P(ohp_stream_spec, chunk_frames); P(ohp_stream_spec, out_fmt); P(ohp_stream_spec, total_frames);
  P(ohp_stream_spec, src_base); P(ohp_stream_spec, dst_base); P(ohp_stream_spec, first_event); P(ohp_stream_spec, driver_block_frames);
  printf("ohp_chunk_info %zu\n", sizeof(ohp_chunk_info));
  printf("ohp_ramp %zu\n", sizeof(ohp_ramp));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "l.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "l")
        subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout
    got = dict(line.rsplit(" ", 1) for line in out.strip().splitlines())
    dts = {"ohp_chunk_desc": abi.CHUNK_DESC, "ohp_ramp_event": abi.RAMP_EVENT, "ohp_stream_spec": abi.STREAM_SPEC,
           "ohp_chunk_info": abi.CHUNK_INFO, "ohp_ramp": abi.RAMP}
    for name, dt in dts.items():
        assert int(got[name]) == dt.itemsize, name
    for key, val in got.items():
        if "." in key:
            t, f = key.split(".")
            assert dts[t].fields[f][1] == int(val), key


def test_ramp_table_matches_oracle(port):
    assert np.array_equal(capi.ramp_table().astype(np.uint32), port.ramp_array)


def test_no_device_is_reported_not_hidden():
    """Without a GPU the library must say so (no CPU fallback)."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.OhpError) as e:
        capi.Context(0)
    assert e.value.status == abi.E_NO_DEVICE


def test_validate_rejects_what_the_reference_asserts_on():
    from util import make_desc
    ok = make_desc(bytes=24, bit_depth=24, channels=2)
    assert capi.validate(ok, 64, 64) == (abi.OK, 0)
    cases = [
        make_desc(bytes=25, bit_depth=24, channels=2),                      # not a whole frame (ProcessorAudioUtils.cpp:45)
        make_desc(bytes=24, bit_depth=20, channels=2),                      # ConstructPcm ASSERT (Msg.cpp:349)
        make_desc(bytes=24, bit_depth=24, channels=0),
        make_desc(bytes=24, bit_depth=24, channels=2, attenuation=128),     # ApplyAttenuation ASSERT(iBitDepth == 16) (Msg.cpp:2741)
        make_desc(bytes=24, bit_depth=24, channels=2, ramp_start=16385),    # Ramp::DoValidate (Msg.cpp:747)
        make_desc(bytes=32, bit_depth=32, channels=2, out_fmt=abi.OUT_PACKED_LE),  # ProcessorPcmSwpEndianPacked 32-bit ASSERTS
        make_desc(bytes=24, bit_depth=24, channels=2, out_fmt=abi.OUT_PACKED_LE, flags=abi.F_SILENCE),  # ProcessSilence ASSERTS
        make_desc(bytes=9216 + 6, bit_depth=24, channels=2),                # larger than a DecodedAudio cell
    ]
    for d in cases:
        assert capi.validate(d, 1 << 20, 1 << 20)[0] == abi.E_INVALID_DESC, d
    assert capi.validate(make_desc(bytes=24, bit_depth=24, channels=2, src_off=60), 64, 64)[0] == abi.E_OUT_OF_RANGE
    assert capi.validate(make_desc(bytes=24, bit_depth=24, channels=2, dst_off=60), 64, 64)[0] == abi.E_OUT_OF_RANGE


def test_headers_are_plain_c99_and_cxx11(tmp_path):
    """The drop-in boundary is a C ABI: every header in include/ compiles alone as strict C99 and as C++11 (no torch, no
    CUDA types in the signatures), so cgo / JNI / ctypes / a C++ adapter can all bind it."""
    import glob
    import subprocess
    for hdr in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = tmp_path / "one.c"
        src.write_text('#include "%s"\nint main(void) { return 0; }\n' % os.path.basename(hdr))
        for cmd in (["gcc", "-std=c99"], ["g++", "-std=c++11", "-x", "c++"]):
            r = subprocess.run(cmd + ["-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c", str(src),
                                      "-o", str(tmp_path / "one.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            assert r.returncode == 0, (hdr, cmd, r.stdout[-2000:])


def test_a_plain_c99_program_links_and_drives_the_boundary(tmp_path):
    """tests/c/abi_client.c: nothing but include/*.h and the two shared libraries, strict C99 -- what a cgo / JNI stub binds.
    The host control plane of one starved stream works from C, the descriptors it produces pass the device calls' own
    validation, and without a GPU every way into the compute path refuses (no fallback)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    capi.cuda_lib(), capi.host_lib()  # built
    pkg = os.path.join(ROOT, "ohpipeline_b200")
    exe = str(tmp_path / "abi_client")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(ROOT, "tests", "c", "abi_client.c"), "-L" + pkg, "-lohp_b200", "-lohp_host", "-Wl,-rpath," + pkg],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-3000:]
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines() if " " in line)
    assert int(out["chunks"]) == 22 and int(out["starved_at_ramp"]) == abi.RAMP_MAX
    if capi.device_count() == 0:
        assert int(out["devices"]) == 0 and "no CPU fallback" in out["refused"]
