"""Drop-in for model/networks/block_extractor/block_extractor.py (same class
names, constructor and call signatures, same shape checks and errors)."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import functional as F_


class BlockExtractorFunction(Function):
    """reference: block_extractor.py:5-42"""

    @staticmethod
    def forward(ctx, source, flow_field, kernel_size):
        assert source.is_contiguous()
        assert flow_field.is_contiguous()
        ctx.save_for_backward(source, flow_field)
        ctx.kernel_size = kernel_size
        return F_.block_extract_fwd(source, flow_field, kernel_size)

    @staticmethod
    def backward(ctx, grad_output):
        source, flow_field = ctx.saved_tensors
        grad_source, grad_flow_field = F_.block_extract_bwd(source, flow_field, grad_output, ctx.kernel_size)
        return grad_source, grad_flow_field, None


class BlockExtractor(Module):
    """reference: block_extractor.py:45-54"""

    def __init__(self, kernel_size=3):
        super(BlockExtractor, self).__init__()
        self.kernel_size = kernel_size

    def forward(self, source, flow_field):
        source_c = source.contiguous()
        flow_field_c = flow_field.contiguous()
        return BlockExtractorFunction.apply(source_c, flow_field_c, self.kernel_size)
