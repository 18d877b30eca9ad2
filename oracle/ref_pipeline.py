"""TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's ExtractorAttn tail (model/networks/base_function.py:804-810) on the
CPU: the reference's own kernel bodies (oracle/_ref, or our restatement when the
reference build is unavailable) for block_extractor / local_attn_reshape, composed
with stock torch CPU ops for softmax / multiply / avg_pool2d and differentiated by
torch autograd -- the same structure as block_extractor.py:5-42 and
local_attn_reshape.py:5-37.  Used by tests/golden/make_golden.py and by bench.py's
`cpu_baseline` and `--impl reference` legs; never by the product path.
"""
import numpy as np
import torch


def make_functions(lib):
    """lib: oracle.oracle.Ref (reference bodies) or oracle.oracle.Oracle (restatement)."""

    class ExtractFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, source, flow, k):
            ctx.save_for_backward(source, flow)
            ctx.k = k
            return torch.from_numpy(lib.block_extract_fwd(source.numpy(), flow.numpy(), k))

        @staticmethod
        def backward(ctx, g):
            s, f = ctx.saved_tensors
            gs, gf = lib.block_extract_bwd(s.numpy(), f.numpy(), np.ascontiguousarray(g.numpy()), ctx.k)
            return torch.from_numpy(gs), torch.from_numpy(gf), None

    class ReshapeFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, k):
            ctx.save_for_backward(x)
            ctx.k = k
            return torch.from_numpy(lib.attn_reshape_fwd(x.numpy(), k))

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            return torch.from_numpy(lib.attn_reshape_bwd(x.numpy(), np.ascontiguousarray(g.numpy()), ctx.k)), None

    return ExtractFn, ReshapeFn


def local_attn_fwd_bwd(lib, source, flow, logits, grad_out, k):
    """numpy in, numpy out: (out, probs, grad_source, grad_flow, grad_logits)."""
    ExtractFn, ReshapeFn = make_functions(lib)
    ts = torch.from_numpy(source).requires_grad_()
    tf = torch.from_numpy(flow).requires_grad_()
    tl = torch.from_numpy(logits).requires_grad_()
    block = ExtractFn.apply(ts, tf, k)                           # base_function.py:805
    probs = torch.softmax(tl, dim=1)                             # nn.Softmax(dim=1), :795,:803
    attn = ReshapeFn.apply(probs, k)                             # :808
    out = torch.nn.functional.avg_pool2d(attn * block, k, k)     # :809
    out.backward(torch.from_numpy(grad_out))
    return (out.detach().numpy(), probs.detach().numpy(), ts.grad.numpy(), tf.grad.numpy(), tl.grad.numpy())
