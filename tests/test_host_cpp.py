"""Runs the C++ reference-style suites of the host mirror (tests/cpp/test_host_api.cpp, built by
__graft_entry__.build()): SuiteRamp / SuiteMsgAudio / SuiteMsgPlayable restated against ohp::media::*, and -- on the
GPU box -- MsgPlayable::Read through BatchPcmReader (C ABI, HOST buffers) checked against the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api")


def run(args):
    r = subprocess.run([EXE] + args, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return r.returncode, r.stdout


def test_cpp_host_mirror_suites():
    rc, out = run([])
    assert rc == 0, out
    assert "PASS" in out


@pytest.mark.gpu
def test_cpp_batch_reader_on_gpu():
    rc, out = run(["--gpu", "--oracle", os.path.join(ROOT, "oracle", "libohp_oracle.so")])
    assert rc == 0, out
    assert "with GPU suite" in out
