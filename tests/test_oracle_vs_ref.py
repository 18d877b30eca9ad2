"""CPU, build container only: the oracle against the reference's own kernel
bodies compiled for the host (oracle/_ref), on fresh random inputs -- bit for bit,
one host thread so the atomics' accumulation order is the thread-index order."""
import numpy as np
import pytest


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1, 8, 32, 32, 32, 32, 3), (2, 3, 7, 9, 7, 9, 4), (2, 4, 9, 11, 6, 5, 5), (1, 2, 5, 5, 5, 5, 2)])
def test_block_extractor(oracle_lib, ref_lib, dt, shape):
    B, C, Hs, Ws, H, W, k = shape
    rng = np.random.default_rng(hash(shape) % 2**32)
    s = rng.standard_normal((B, C, Hs, Ws)).astype(dt)
    f = rng.uniform(-1.5 * W, 1.5 * W, (B, 2, H, W)).astype(dt)
    a, b = oracle_lib.block_extract_fwd(s, f, k), ref_lib.block_extract_fwd(s, f, k)
    assert np.array_equal(a, b)
    g = rng.standard_normal(a.shape).astype(dt)
    for x, y in zip(oracle_lib.block_extract_bwd(s, f, g, k), ref_lib.block_extract_bwd(s, f, g, k)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("cfg", [(2, 3, 8, 9, 8, 9, 2, 1, 5.0, 4.0), (1, 4, 10, 12, 7, 6, 4, 1, 2.0, 12.0), (1, 2, 9, 9, 9, 9, 4, 2, 2.0, 3.0), (1, 2, 6, 6, 6, 6, 2, 1, 0.0, 2.0)])
def test_resample2d(oracle_lib, ref_lib, dt, cfg):
    B, C, Hi, Wi, H, W, ks, dil, sig, amp = cfg
    rng = np.random.default_rng(7)
    a1 = rng.standard_normal((B, C, Hi, Wi)).astype(dt)
    in2 = np.concatenate([rng.uniform(-amp, amp, (B, 2, H, W)), np.full((B, 1, H, W), sig)], 1).astype(dt)
    oa, ob = oracle_lib.resample2d_fwd(a1, in2, ks, dil), ref_lib.resample2d_fwd(a1, in2, ks, dil)
    assert np.array_equal(oa, ob, equal_nan=True)
    g = rng.standard_normal(oa.shape).astype(dt)
    for x, y in zip(oracle_lib.resample2d_bwd(a1, in2, g, ks, dil), ref_lib.resample2d_bwd(a1, in2, g, ks, dil)):
        assert np.array_equal(x, y, equal_nan=True)


def test_reshape(oracle_lib, ref_lib):
    rng = np.random.default_rng(3)
    for k in (2, 3, 4, 5):
        x = rng.standard_normal((2, k * k, 5, 7)).astype(np.float32)
        a = oracle_lib.attn_reshape_fwd(x, k)
        assert np.array_equal(a, ref_lib.attn_reshape_fwd(x, k))
        g = rng.standard_normal(a.shape).astype(np.float32)
        assert np.array_equal(oracle_lib.attn_reshape_bwd(x, g, k), ref_lib.attn_reshape_bwd(x, g, k))
