"""TEST INFRASTRUCTURE ONLY -- platform-independent deterministic tensors for the golden vectors.

The golden fixtures store only *outputs*; inputs and weights are regenerated wherever the tests
run from this integer hash (no dependence on torch / numpy RNG streams, which may differ between
library versions).  Values are uniform; ``uniform_linear`` mirrors torch's default nn.Linear
initialisation range U(-1/sqrt(fan_in), 1/sqrt(fan_in)) used by the reference's layers
(models/attention_processor.py:51-56, models/adapters.py:14-28).
"""
import numpy as np
import torch

_M32 = np.uint64(0xFFFFFFFF)


def _hash_u32(idx: np.ndarray, seed: int) -> np.ndarray:
    """Murmur3-style finaliser over (index, seed) -> uint32, computed in uint64 with masking."""
    x = (idx.astype(np.uint64) + np.uint64((seed * 0x9E3779B1) & 0xFFFFFFFF)) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def uniform(shape, seed: int, lo: float = -1.0, hi: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    h = _hash_u32(np.arange(n, dtype=np.uint64), seed)
    # 24 random bits -> exactly representable in fp32, identical on every platform
    u = (h >> np.uint64(8)).astype(np.float64) / float(1 << 24)
    v = lo + (hi - lo) * u
    return torch.from_numpy(v.reshape(shape)).to(dtype)


def unit_variance(shape, seed: int, dtype=torch.float32) -> torch.Tensor:
    """Zero-mean, unit-variance activations (uniform on +-sqrt(3))."""
    a = 3.0 ** 0.5
    return uniform(shape, seed, -a, a, dtype)


def uniform_linear(out_features: int, in_features: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    b = 1.0 / (in_features ** 0.5)
    return uniform((out_features, in_features), seed, -b, b, dtype)


def uniform_bias(out_features: int, in_features: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    b = 1.0 / (in_features ** 0.5)
    return uniform((out_features,), seed, -b, b, dtype)
