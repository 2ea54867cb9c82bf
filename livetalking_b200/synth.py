"""Synthetic (random-init) wav2lip256 weights and avatar assets in the reference's formats, for benchmarking and
bring-up when no checkpoint / avatar directory is available (there is no network in the build environment).

The state_dict uses the reference checkpoint's key scheme (avatars/wav2lip_avatar.py:59-70); BN layers are
identity (gamma=1, beta=0, mean=0, var=1) and residual branches are damped so activations stay in fp16 range.
Timing does not depend on the weight values (no data-dependent control flow anywhere in the path)."""
from __future__ import annotations

import numpy as np

from .w2l_pack import layer_table

_RES = {  # prefixes of residual Conv2d blocks (wav2lip_v2.py: residual=True)
}


def random_state_dict(seed: int = 0):
    rng = np.random.default_rng(seed)
    sd = {}
    for prefix, kind, ci, co, k in layer_table():
        shape = (co, ci, k, k) if kind == "c" else (ci, co, k, k)
        fan_in = ci * k * k if kind == "c" else ci * k * k / 4.0
        residual = (kind == "c" and ci == co and k == 3 and not prefix.endswith(".0") and "output_block" not in prefix)
        std = np.sqrt(2.0 / fan_in) * (0.3 if residual else 1.0)
        sd[f"{prefix}.conv_block.0.weight"] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std))
        sd[f"{prefix}.conv_block.0.bias"] = (rng.standard_normal(co, dtype=np.float32) * np.float32(0.02))
        sd[f"{prefix}.conv_block.1.weight"] = np.ones(co, np.float32)
        sd[f"{prefix}.conv_block.1.bias"] = np.zeros(co, np.float32)
        sd[f"{prefix}.conv_block.1.running_mean"] = np.zeros(co, np.float32)
        sd[f"{prefix}.conv_block.1.running_var"] = np.ones(co, np.float32)
    sd["output_block.1.weight"] = rng.standard_normal((3, 32, 1, 1), dtype=np.float32) * np.float32(0.3)
    sd["output_block.1.bias"] = np.zeros(3, np.float32)
    return sd


def synthetic_avatar(n: int = 64, H: int = 720, W: int = 1280, bbox=(200, 520, 480, 800), seed: int = 0):
    """seeded faces (n,256,256,3) u8, full frames (n,H,W,3) u8, coords (n,4) = (y1,y2,x1,x2)."""
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (n, 16, 16, 3)).astype(np.float32)
    faces = np.clip(np.kron(low, np.ones((1, 16, 16, 1), np.float32)) + rng.integers(-8, 9, (n, 256, 256, 3)), 0, 255).astype(np.uint8)
    base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames = np.empty((n, H, W, 3), np.uint8)
    for i in range(n):
        frames[i] = np.roll(base, i * 3, axis=1)
    coords = np.tile(np.asarray(bbox, np.int32), (n, 1))
    return faces, frames, coords


def sine_audio(seconds: float = 60.0, freq: float = 440.0, amp: float = 0.5, sr: int = 16000) -> np.ndarray:
    """BASELINE.json synthetic audio: amp*sin(2 pi f t), 16 kHz float32."""
    t = np.arange(int(seconds * sr), dtype=np.float64) / sr
    return (amp * np.sin(2 * np.pi * freq * t)).astype(np.float32)
