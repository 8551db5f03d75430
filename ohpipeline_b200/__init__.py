"""ohpipeline_b200 -- B200-native batch implementation of ohPipeline's decoded-PCM ramp + format-convert path.

Layout:
  csrc/     CUDA kernels (sm_100a) and the C ABI of include/ohp_b200.h        -> libohp_b200.so
  host/     C++ mirror of the reference's Ramp / MsgAudio / MsgPlayable /
            IPcmProcessor interface, the stage chain and the schedule C ABI  -> libohp_host.so
  abi.py    numpy dtypes of the ABI structs;  capi.py  ctypes binding;  workloads.py  synthetic configs
"""
from . import abi  # noqa: F401

__all__ = ["abi", "capi", "workloads"]
