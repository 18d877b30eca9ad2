"""Drop-in for model/networks/local_attn_reshape/local_attn_reshape.py."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import functional as F_


class LocalAttnReshapeFunction(Function):
    """reference: local_attn_reshape.py:5-37"""

    @staticmethod
    def forward(ctx, inputs, kernel_size):
        assert inputs.is_contiguous()
        ctx.kernel_size = kernel_size
        return F_.attn_reshape_fwd(inputs, kernel_size)

    @staticmethod
    def backward(ctx, grad_output):
        return F_.attn_reshape_bwd(grad_output, ctx.kernel_size), None


class LocalAttnReshape(Module):
    """reference: local_attn_reshape.py:40-46 (kernel_size is a CALL argument)"""

    def __init__(self):
        super(LocalAttnReshape, self).__init__()

    def forward(self, inputs, kernel_size=3):
        inputs_c = inputs.contiguous()
        return LocalAttnReshapeFunction.apply(inputs_c, kernel_size)
