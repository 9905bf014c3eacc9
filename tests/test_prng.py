"""Threefry-2x32 known-answer vectors (Random123 kat_vectors, also used by jax's
own random_test.py) and the well-known jax.random values for PRNGKey(0)."""
import numpy as np

from rsr_mjx_b200 import prng


def test_threefry_kat():
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, exp in kat:
        y = prng.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(y[0]), int(y[1])) == exp


def test_split_matches_jax_documented_values():
    # jax.random.split(jax.random.PRNGKey(0)) under the default threefry impl
    k = prng.split(prng.PRNGKey(0))
    assert k.tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_uniform_range_and_batching():
    keys = prng.split(prng.PRNGKey(7), 64)
    u = prng.uniform(keys, (5,), -0.01, 0.01)
    assert u.shape == (64, 5) and u.dtype == np.float32
    assert (u >= -0.01).all() and (u < 0.01).all()
    # batched == per-key
    for i in (0, 13, 63):
        np.testing.assert_array_equal(u[i], prng.uniform(keys[i], (5,), -0.01, 0.01))
    # odd sizes pad the counter array like jax does
    u3 = prng.uniform(prng.PRNGKey(3), (3,))
    assert u3.shape == (3,) and (u3 >= 0).all() and (u3 < 1).all()
    # degenerate range (the reference draws z from [0.82, 0.82])
    z = prng.uniform(prng.PRNGKey(1), (3,), np.float32([0.5, -0.005, 0.82]), np.float32([0.51, 0.005, 0.82]))
    assert z[2] == np.float32(0.82)
