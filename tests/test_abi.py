"""CPU: the C-ABI library loads, exports every symbol include/gfla_warp.h declares,
and rejects bad arguments before touching a device (no compute is launched here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def so():
    import __graft_entry__ as ge
    ge.build_cuda()
    import gfla_b200
    return gfla_b200._lib.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "gfla_warp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*|unsigned long long|long long)\s+(gfla_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        decls[m.group(1)] = 0 if args == ["void"] else len(args)
    return decls


def test_header_symbols_exported(so):
    import gfla_b200
    decls = _declared()
    assert len(decls) >= 11
    for name, nargs in decls.items():
        assert hasattr(so, name), f"{name} declared in gfla_warp.h but not exported"
        if name in gfla_b200._lib.SIGNATURES:
            assert len(gfla_b200._lib.SIGNATURES[name]) == nargs, name
    for name in gfla_b200._lib.SIGNATURES:
        assert name in decls, f"{name} bound in _lib.py but not declared in the header"


def test_abi_version_and_error_strings(so):
    assert so.gfla_abi_version() == 1
    assert b"NULL" in so.gfla_error_string(-1)
    assert b"shape" in so.gfla_error_string(-2)
    assert so.gfla_error_string(0) == b"ok"


def test_argument_validation_needs_no_device(so):
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    # NULL pointer
    assert so.gfla_block_extract_fwd(None, p, p, 1, 1, 2, 2, 2, 2, 3, 0, 0, None) == -1
    # kernel_size out of range / non-positive dims
    assert so.gfla_block_extract_fwd(p, p, p, 1, 1, 2, 2, 2, 2, 0, 0, 0, None) == -2
    assert so.gfla_block_extract_fwd(p, p, p, 0, 1, 2, 2, 2, 2, 3, 0, 0, None) == -2
    # dtype: unknown code; fp32 data with fp64 flow; resample2d is float/double only
    assert so.gfla_block_extract_fwd(p, p, p, 1, 1, 2, 2, 2, 2, 3, 7, 7, None) == -3
    assert so.gfla_block_extract_fwd(p, p, p, 1, 1, 2, 2, 2, 2, 3, 0, 1, None) == -3
    assert so.gfla_resample2d_fwd(p, p, p, 1, 1, 2, 2, 2, 2, 2, 1, 2, None) == -3
    # misaligned pointer
    assert so.gfla_attn_reshape_fwd(p + 2, p, 1, 2, 2, 3, 0, None) == -4
    # unknown algo / tile kernel asked for a dtype it cannot serve
    assert so.gfla_local_attn_fwd(p, p, p, p, None, 1, 1, 2, 2, 2, 2, 3, 0, 0, 0, 9, None) == -5
    assert so.gfla_local_attn_fwd(p, p, p, p, None, 1, 1, 2, 2, 2, 2, 3, 1, 1, 0, 2, None) == -5
    # unknown layout code
    assert so.gfla_local_attn_fwd(p, p, p, p, None, 1, 1, 2, 2, 2, 2, 3, 0, 0, 5, 0, None) == -2


def test_cpu_tensors_raise_like_the_reference():
    import torch
    import gfla_b200
    with pytest.raises(NotImplementedError):      # block_extractor.py:23-24
        gfla_b200.BlockExtractor(3)(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(NotImplementedError):      # local_attn_reshape.py:20-21
        gfla_b200.LocalAttnReshape()(torch.zeros(1, 9, 4, 4), 3)
    with pytest.raises(AssertionError):           # block_extractor.py:16  (df == 2)
        gfla_b200.BlockExtractorFunction.apply(torch.zeros(1, 2, 4, 4), torch.zeros(1, 3, 4, 4), 3)
    with pytest.raises(AssertionError):           # local_attn_reshape.py:13 (ds == k*k)
        gfla_b200.LocalAttnReshapeFunction.apply(torch.zeros(1, 8, 4, 4), 3)


def test_legacy_module_names():
    import gfla_b200
    gfla_b200.compat.install()
    import block_extractor_cuda, local_attn_reshape_cuda, resample2d_cuda  # noqa: E401
    for m in (block_extractor_cuda, local_attn_reshape_cuda, resample2d_cuda):
        assert callable(m.forward) and callable(m.backward)
    from model.networks.block_extractor.block_extractor import BlockExtractor
    from model.networks.local_attn_reshape.local_attn_reshape import LocalAttnReshape
    from model.networks.resample2d_package.resample2d import Resample2d
    assert BlockExtractor(5).kernel_size == 5 and LocalAttnReshape() is not None
    assert Resample2d(4, 1, sigma=2).sigma == 2.0       # constructible without a GPU (unlike resample2d.py:47)


def test_extractor_attn_state_dict_keys():
    """checkpoint compatibility: same parameter names as base_function.py:799-803"""
    import gfla_b200
    m = gfla_b200.ExtractorAttn(8, 3, softmax=True)
    assert sorted(m.state_dict()) == ["fully_connect_layer.0.bias", "fully_connect_layer.0.weight",
                                      "fully_connect_layer.2.bias", "fully_connect_layer.2.weight"]
    assert tuple(m.state_dict()["fully_connect_layer.0.weight"].shape) == (128, 16, 3, 3)
    assert tuple(m.state_dict()["fully_connect_layer.2.weight"].shape) == (9, 128, 1, 1)
