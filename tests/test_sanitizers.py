"""Host code under AddressSanitizer + UndefinedBehaviorSanitizer (the reference runs its tests under valgrind,
SURVEY 4): the container parser on truncated and mutated files, and the schedule model + the class-free walk -- the
source the GPU kernels compile -- on randomized workloads, stream by stream, with their results compared on the way.
Inputs live in exact-size heap blocks so that any over-read is a report.  CPU only."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from ohpipeline_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ohpipeline_b200", "host")
FLAGS = ["-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
         "-I" + os.path.join(ROOT, "include")]


def build(tmp_path, name, sources):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / name)
    r = subprocess.run(["g++"] + FLAGS + ["-o", exe, os.path.join(ROOT, "tests", "sanitize", name + ".cpp")] + sources,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and "sanitize" in r.stdout and "cannot find" in r.stdout:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stdout[-3000:]
    return exe


def run(exe, path):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([exe, path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "no sanitizer report" in r.stdout, r.stdout[-3000:]
    return r.stdout


def test_container_parser_under_sanitizers(tmp_path):
    from test_container import aiff_bytes, wav_bytes
    rng = np.random.default_rng(12)
    seeds = [wav_bytes(44100, 16, 2, 50)[0], wav_bytes(96000, 24, 6, 20)[0], aiff_bytes(48000, 24, 2, 40)[0],
             aiff_bytes(44100, 16, 2, 40, sowt=True)[0], aiff_bytes(22050, 8, 1, 33)[0]]
    path = str(tmp_path / "containers.bin")
    with open(path, "wb") as out:
        for data in seeds:
            for cut in range(len(data) + 1):
                out.write(struct.pack("<I", cut) + data[:cut])
            for _ in range(3000):
                b = bytearray(data[:int(rng.integers(0, len(data) + 1))])
                for _ in range(int(rng.integers(1, 8))):
                    if b:
                        b[int(rng.integers(0, min(len(b), 80)))] = int(rng.integers(0, 256))
                out.write(struct.pack("<I", len(b)) + bytes(b))
    exe = build(tmp_path, "container_fuzz", [os.path.join(HOST, "container.cpp")])
    assert "usable" in run(exe, path)


def test_schedule_model_and_walk_under_sanitizers(tmp_path):
    path = str(tmp_path / "workloads.bin")
    with open(path, "wb") as out:
        def dump(w):
            out.write(struct.pack("<II", len(w.streams), len(w.events)) + w.streams.tobytes() + w.events.tobytes())
        for seed in range(700, 706):
            dump(workloads.mixed(n_streams=80, seed=seed, max_frames=5000))
            dump(workloads.steady_edges(seed, n_streams=80))
            dump(workloads.config4(n_streams=40, seconds=0.15, seed=seed))
            dump(workloads.elements(seed, n_streams=40))   # element state machines; starvations -> ohp_flywheel_plan
        dump(workloads.all_rates(seconds=0.15))
        dump(workloads.config3(n_streams=30, seconds=1.0))
    exe = build(tmp_path, "schedule_fuzz", [os.path.join(HOST, "schedule.cpp"), os.path.join(HOST, "msg_model.cpp")])
    out = run(exe, path)
    assert "chunks" in out
    planned = int(out.split(" starvations planned")[0].split()[-1])
    assert planned >= 50, out


def test_bulk_step_covers_the_steady_stretches(tmp_path):
    """A performance property of the GPU schedule builder that can be checked on the CPU: on the BASELINE configs nearly
    every message goes through the bulk step (32 at a time); only what an event, a ramp end or a split touches is
    walked one at a time."""
    cases = [("config2", workloads.config2(n_streams=4, seconds=10.0), 0.99, 28.0),
             ("config5", workloads.config5(n_streams=8, seconds=1.0), 0.97, 20.0),
             ("config3", workloads.config3(n_streams=64, seconds=1.0), 0.80, 4.0),
             ("config4", workloads.config4(n_streams=256, seconds=0.25), 0.75, 4.0),
             ("all_rates", workloads.all_rates(), 0.80, 4.0)]
    aiff = workloads.config2(n_streams=4, seconds=10.0)
    aiff.streams["sample_rate"] = 44100; aiff.streams["bit_depth"] = 16; aiff.streams["chunk_frames"] = 220
    aiff.streams["total_frames"] = 441000; aiff.streams["codec_read_frames"] = 9216 // 4
    cases.append(("aiff-born", aiff, 0.99, 25.0))
    path = str(tmp_path / "cases.bin")
    with open(path, "wb") as out:
        for _, w, _, _ in cases:
            out.write(struct.pack("<II", len(w.streams), len(w.events)) + w.streams.tobytes() + w.events.tobytes())
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "walk_stats")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(ROOT, "tests", "sanitize", "walk_stats.cpp")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    lines = subprocess.run([exe, path], stdout=subprocess.PIPE, text=True, check=True).stdout.strip().splitlines()
    assert len(lines) == len(cases)
    for (name, _, min_share, min_run), line in zip(cases, lines):
        chunks, calls, bulk, general = (int(x) for x in line.split())
        assert chunks > 0 and calls > 0, name
        share = bulk / (bulk + general)
        assert share >= min_share, (name, share)
        assert bulk / calls >= min_run, (name, bulk / calls)


def test_cpp_host_mirror_suite_under_sanitizers(tmp_path):
    """The reference-style C++ suites of the host mirror (tests/cpp/test_host_api.cpp, CPU part: ramp algebra and the
    message model with its ref-counted, recycled messages) under ASan + UBSan + leak check."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    pkg = os.path.join(ROOT, "ohpipeline_b200")
    exe = str(tmp_path / "test_host_api_san")
    r = subprocess.run(["g++"] + FLAGS + ["-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
                                          os.path.join(HOST, "msg_model.cpp"), "-L" + pkg, "-lohp_b200", "-Wl,-rpath," + pkg, "-ldl"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and "sanitize" in r.stdout and "cannot find" in r.stdout:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stdout[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout and "0 failures" in r.stdout, r.stdout[-3000:]
