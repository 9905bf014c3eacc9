"""NumPy restatement of jax.random's default threefry2x32 PRNG (jax==0.4.29,
`jax_threefry_partitionable=False`): PRNGKey, split, uniform.

The reference draws every reset / domain-randomisation sample from
``jax.random.split`` + ``jax.random.uniform`` (reference test/airbot.py:104-133,
ppo_train/airbot_training/domain_randomize.py:36-69).  jax is not installable
here, so bit-equality with jax is *unverified*; the Threefry-2x32 core is checked
against the Random123 known-answer vectors in tests/test_prng.py.

All functions are vectorised over leading batch dimensions of ``key``.
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds. All args uint32 arrays (broadcastable)."""
    with np.errstate(over="ignore"):
        k0, k1 = np.asarray(k0, _U32), np.asarray(k1, _U32)
        x0, x1 = np.array(x0, _U32), np.array(x1, _U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def _threefry_2x32_counts(key, n):
    """jax._src.prng.threefry_2x32(key, iota(n)) for key[..., 2] -> [..., n]."""
    key = np.asarray(key, _U32)
    odd = n % 2
    cnt = np.arange(n + odd, dtype=_U32)
    h = (n + odd) // 2
    x0, x1 = cnt[:h], cnt[h:]
    y0, y1 = threefry2x32(key[..., 0:1], key[..., 1:2], x0, x1)
    out = np.concatenate([y0, y1], axis=-1)
    return out[..., :n]


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key, num: int = 2) -> np.ndarray:
    """key[..., 2] -> keys[..., num, 2]"""
    bits = _threefry_2x32_counts(key, 2 * num)
    return bits.reshape(bits.shape[:-1] + (num, 2))


def fold_in(key, data: int) -> np.ndarray:
    """jax.random.fold_in: threefry_2x32(key, threefry_seed(data)), threefry_seed(uint32 d) = [0, d] (unverified
    against jax, like the rest of this module)"""
    key = np.asarray(key, _U32)
    y0, y1 = threefry2x32(key[..., 0], key[..., 1], _U32(0), _U32(int(data) & 0xFFFFFFFF))
    return np.stack([y0, y1], axis=-1).astype(_U32)


def random_bits(key, n: int) -> np.ndarray:
    return _threefry_2x32_counts(key, n)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    """float32 U[minval, maxval) of `shape` per key; key[..., 2] -> [..., *shape]."""
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    bits = random_bits(key, n)
    fb = (bits >> _U32(9)) | _U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    f = f.reshape(f.shape[:-1] + shape)
    lo = np.asarray(minval, np.float32)
    hi = np.asarray(maxval, np.float32)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
