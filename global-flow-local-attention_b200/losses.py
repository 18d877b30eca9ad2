"""Flow-regularisation loss of the reference, without the warping ops.

Reference: model/networks/external_function.py:12-77 (``MultiAffineRegularizationLoss``,
``AffineRegularizationLoss``).  Same constructors, same call signatures, same values -- SURVEY.md section 8 row f3.

The reference evaluates, for every kz x kz window p of the sampling grid (x or y component),

    results  = conv2d(grid, M)                      M = K^T K  (kz^2 x kz^2),  K = A (A^T A)^-1 A^T - I
    kernels  = LocalAttnReshape(results, kz)        [b, 1, kz h', kz w']
    grid_H   = BlockExtractor(kz)(grid, const flow int(kz/2))
    loss     = mean(avg_pool2d(grid_H * kernels, kz, kz)) * kz^2

With the constant INTEGER flow int(kz/2) every bilinear tap of the extractor lands exactly on a pixel (weights
1 and 0, block_extractor_kernel.cu:62-82), and because the flow field is (kz-1) smaller than the grid no tap is
clamped: ``grid_H`` is just the kz x kz patch around every window, laid out the way ``LocalAttnReshape`` lays
out ``results``.  So  avg_pool(grid_H * kernels) * kz^2  ==  sum_q results[q] * patch[q]  ==  p^T M p, and the
two inflated [b, 1, kz h', kz w'] tensors, their memsets and the two custom-op launches per component (and
their atomics in the backward) disappear: what is left is the kz^2-channel convolution, an ``unfold`` and a
reduction, all differentiable by autograd.  Nothing here needs the CUDA library, so unlike the reference's
class it also runs on CPU tensors (tests/test_losses.py checks it there against the literal composition
evaluated with the oracle's block_extractor / local_attn_reshape).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def affine_residual_kernel(kz: int) -> np.ndarray:
    """M = K^T K with K = A (A^T A)^-1 A^T - I, A = rows (i, j, 1) of the window positions q = i*kz + j
    (external_function.py:40-46): p^T M p is the squared residual of the best affine fit to the window p."""
    i, j = np.meshgrid(np.arange(kz), np.arange(kz), indexing="ij")
    a = np.stack([i.ravel(), j.ravel(), np.ones(kz * kz)], axis=1).astype(np.float64)
    k = a @ np.linalg.inv(a.T @ a) @ a.T - np.eye(kz * kz)
    return k.T @ k


class AffineRegularizationLoss(nn.Module):
    """Drop-in for external_function.py:31-77: ``loss = AffineRegularizationLoss(kz)(flow_field)``."""

    def __init__(self, kz):
        super(AffineRegularizationLoss, self).__init__()
        self.kz = kz
        # same tensor as the reference's ``self.kernel``: [kz^2, 1, kz, kz], float64 until ``type_as`` at call time
        self.kernel = torch.from_numpy(affine_residual_kernel(kz)).view(kz * kz, 1, kz, kz)

    def __call__(self, flow_fields):
        grid = self.flow2grid(flow_fields)
        weights = self.kernel.type_as(flow_fields)
        return self.calculate_loss(grid[:, 0:1], weights) + self.calculate_loss(grid[:, 1:2], weights)

    def calculate_loss(self, grid, weights):
        kz = self.kz
        results = F.conv2d(grid, weights)                                  # [b, kz^2, h', w']  = M p per window
        b, c, h, w = results.size()
        patches = F.unfold(grid, kz).view(b, c, h, w)                      # p itself, channel q = i*kz + j
        return torch.mean(torch.sum(results * patches, dim=1, keepdim=True))

    def flow2grid(self, flow_field):
        b, c, h, w = flow_field.size()
        x = torch.arange(w, device=flow_field.device).view(1, -1).expand(h, -1).type_as(flow_field).float()
        y = torch.arange(h, device=flow_field.device).view(-1, 1).expand(-1, w).type_as(flow_field).float()
        grid = torch.stack([x, y], dim=0).unsqueeze(0).expand(b, -1, -1, -1)
        return flow_field + grid


class MultiAffineRegularizationLoss(nn.Module):
    """Drop-in for external_function.py:12-28: one ``AffineRegularizationLoss`` per attention level, applied to
    the flow fields in descending layer order."""

    def __init__(self, kz_dic):
        super(MultiAffineRegularizationLoss, self).__init__()
        self.kz_dic = kz_dic
        self.method_dic = {key: AffineRegularizationLoss(kz_dic[key]) for key in kz_dic}
        self.layers = sorted(kz_dic, reverse=True)

    def __call__(self, flow_fields):
        loss = 0
        for i in range(len(flow_fields)):
            loss += self.method_dic[self.layers[i]](flow_fields[i])
        return loss
