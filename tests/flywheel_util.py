"""Shared inputs for the flywheel tests: training blocks in FlywheelInput's layout (planar, 32-bit big-endian)."""
import numpy as np

from ohpipeline_b200 import abi

# (rate, channels, bit depth) accepted by RampGenerator's fixed buffers (ohp_flywheel_validate)
SHAPES = [(44100, 2, 16), (48000, 2, 24), (192000, 2, 24), (96000, 6, 32), (384000, 2, 8), (176400, 1, 16),
          (88200, 8, 24), (352800, 2, 32), (32000, 3, 16), (8000, 2, 8)]


def train_frames(rate):
    return abi.FLYWHEEL_TRAINING_JIFFIES // abi.jiffies_per_sample(rate)


def training_block(rate, channels, kind, seed):
    """kind: 'tone' (what the predictor is for), 'noise' (full-range: every accumulator wraps), 'dc', 'zero',
    'step', 'max'."""
    rng = np.random.default_rng(seed)
    t = np.arange(train_frames(rate))
    planes = []
    for c in range(channels):
        if kind == "tone":
            x = 0.4 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t / rate + 0.3 * c) * 2**31 + rng.normal(0, 2**18, t.size)
        elif kind == "noise":
            x = rng.integers(-2**31, 2**31, t.size).astype(np.float64)
        elif kind == "dc":
            x = np.full(t.size, (c + 1) * 1.1e8)
        elif kind == "zero":
            x = np.zeros(t.size)
        elif kind == "step":
            x = np.where(t < t.size // 2, -1.5e9, 1.5e9)
        else:
            x = np.where(t % 2 == 0, 2.0**31 - 1, -2.0**31)
        planes.append(np.clip(x, -2**31, 2**31 - 1).astype(np.int64).astype(np.int32).astype(">i4").view(np.uint8))
    return np.concatenate(planes)
