"""The reference's OWN unit tests, run against the oracle's build of the reference.

oracle/_ref links Msg.cpp, Ramper.cpp, Muter.cpp, VolumeRamper.cpp, StarvationRamper.cpp, FlywheelRamper.cpp,
DecodedAudioAggregator.cpp ... unmodified, but behind this repo's ohNet shim (oracle/shim: buffers, threads, semaphores,
functors, FIFOs).  Everything pinned "against the reference" in this repo is pinned against THAT build, so the build itself is
held against what the reference's authors check: Media/Tests/TestMsg.cpp (allocator, Ramp, RampApplicator through
MsgPlayable::Read, MsgAudio Split / SetRamp, queues, reservoirs ...), TestRamper.cpp, TestMuter.cpp, TestVolumeRamper.cpp,
TestStarvationRamper.cpp, TestFlywheelRamper.cpp, TestDecodedAudioAggregator.cpp, TestSkipper.cpp, TestWaiter.cpp and
TestVariableDelay.cpp, compiled unmodified
(oracle/Makefile, ref_suites; OpenHome/Private/TestFramework.h and SuiteUnitTest.h are shim headers written here) and run by
oracle/ref_suites_main.cpp.  CPU only; skipped where oracle/_ref did not travel."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_suites")

# suite description -> the least number of TEST()s it must have passed (what it ran when this was written; a shim that
# silently skipped tests would show here).  StarvationRamper's count depends on how its two threads interleave (8403 .. 8485).
EXPECT = {"Ramp tests": 5718, "Basic MsgPlayable tests": 8860, "Basic MsgAudio tests": 3501, "Ramper": 73, "Muter": 153,
          "VolumeRamper": 254, "StarvationRamper": 8000, "SuiteFlywheelRamper": 80, "SuiteDecodedAudioAggregator": 110,
          "MsgQueue tests": 41, "MsgReservoir tests": 100, "Allocator tests": 143,
          # three more elements that ramp with the same Split + SetRamp idiom
          "Skipper": 217, "SuiteWaiter": 438, "VariableDelayLeft": 225, "VariableDelayRight": 51}


def test_the_references_own_suites_pass_on_the_oracle_build(_built):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ref_suites not present (needs /root/reference at build time)")
    # TestStarvationRamper.cpp and TestMuter.cpp give their second thread "a short wait" here and there (Thread::Sleep(50)
    # before looking at what it did): on a loaded machine that can be too short (seen: TestStarvationRamper.cpp:462, one
    # run in 25 with all cores busy), which says nothing about the code under test -- a run that fails is repeated, twice
    # at most
    for attempt in range(3):
        r = subprocess.run([EXE, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        if r.returncode == 0:
            break
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    assert "FAILURE" not in r.stdout, tail
    suites = {}
    for name, passed, failed in re.findall(r"^suite: (.*): (\d+) passed, (\d+) failed$", r.stdout, flags=re.M):
        assert int(failed) == 0, (name, tail)
        suites[name] = max(suites.get(name, 0), int(passed))
    for name, least in EXPECT.items():
        assert suites.get(name, 0) >= least, (name, suites.get(name), tail)
    total = re.search(r"^total: (\d+) passed, 0 failed$", r.stdout, flags=re.M)
    assert total and int(total.group(1)) >= 38500, tail
