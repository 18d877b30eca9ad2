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
