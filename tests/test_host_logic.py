"""CPU: host-side identities the Python layer relies on (checked with the oracle, no GPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_target_half_of_fc_conv_is_a_plain_convolution(oracle_lib, k):
    """ExtractorAttn._logits: conv(kernel k, stride k) over BlockExtractor(target, zero flow) == stride-1 conv of
    `target` with replicate padding (k//2, k-1-k//2) -- exactly (same products, zero-flow taps have weights 1/0)."""
    torch.manual_seed(k)
    B, C, H, W = 2, 3, 7, 6
    t = torch.randn(B, C, H, W)
    w, bias = torch.randn(5, C, k, k), torch.randn(5)
    blk = torch.from_numpy(oracle_lib.block_extract_fwd(t.numpy(), np.zeros((B, 2, H, W), np.float32), k))
    a = F.conv2d(blk, w, bias, stride=k)
    lo, hi = k // 2, k - 1 - k // 2
    b = F.conv2d(F.pad(t, (lo, hi, lo, hi), mode="replicate"), w, bias)
    assert torch.equal(a, b)


def test_feature_layout_detection():
    import gfla_b200
    from gfla_b200 import _lib, functional as F_
    x = torch.zeros(2, 8, 4, 4)
    assert F_._feature_layout(x) == _lib.GFLA_NCHW
    assert F_._feature_layout(x.contiguous(memory_format=torch.channels_last)) == _lib.GFLA_NHWC
    with pytest.raises(AssertionError):
        F_._feature_layout(x[:, ::2])


def test_fastdiv_multiply_shift_is_exact():
    """tc_common.cuh `FastDiv` (group index -> (column, row, sample) in the fused backward): q = umulhi(n, mul) >> shr with
    p = 31 + ceil(log2 d), mul = ceil(2^p / d) (always < 2^32), shr = p - 32 must equal n // d for every 0 <= n < 2^31.
    The formula is restated here and checked for the divisors the kernels can meet (group columns / rows) at boundary and random n."""
    import random

    def make(d):
        p = 31 + (d - 1).bit_length()                 # 32 - clz(d - 1) == bit_length(d - 1)
        m = ((1 << p) + d - 1) // d
        assert m < (1 << 32), d                       # the device keeps it in a uint32_t
        return m, p - 32

    rng = random.Random(7)
    divisors = list(range(2, 70)) + [rng.randrange(70, 1 << 16) for _ in range(300)] + [255, 256, 257, 4095, 4096, 65535, 65536, (1 << 20) + 3]
    for d in divisors:
        mul, shr = make(d)
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - d] + [rng.randrange(0, 1 << 31) for _ in range(200)]:
            assert ((n * mul) >> 32) >> shr == n // d, (n, d)


def test_resample2d_cosine_has_no_cpu_path():
    """like the reference's ops (block_extractor.py:23-24): CPU tensors raise instead of silently running somewhere else"""
    import gfla_b200
    x, t = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    flow = torch.zeros(1, 2, 8, 8)
    with pytest.raises(NotImplementedError):
        gfla_b200.Resample2dCosine(4, 1, sigma=2)(x, flow, t)
    with pytest.raises(NotImplementedError):
        gfla_b200.Resample2d(4, 1, sigma=2)(x, flow)
