"""Losses of the reference that sit on the warping ops: the flow regularisation WITHOUT the ops (row f3) and the sampling
correctness loss on the fused resample -> cosine op (row f4, at the end of this file).

Flow-regularisation loss of the reference, without the warping ops.

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
import math

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


class PerceptualCorrectness(nn.Module):
    """Drop-in for external_function.py:222-284 (the "sampling correctness" loss), SURVEY.md section 8 row f4.

    Same call signature and value.  ``calculate_loss`` differs in ONE step: the reference warps ``source_vgg`` with
    ``Resample2d(4, 1, sigma=2)``, writes the warped features, and reads them back for ``F.cosine_similarity`` against the
    target features (:275-279); here that pair is the fused op ``Resample2dCosine`` -- one kernel forward, one backward, no
    warped tensor, and in the backward no gradient work for the VGG features (they are functions of data).  The N x N
    correlation (:262-271) stays a cuBLAS ``bmm``.

    ``vgg``: the feature extractor (a module returning ``{'relu1_1': ..., ...}``).  ``None`` builds the reference's own ``VGG19``
    (needs ``compat.install(reference_root=...)`` and torchvision's pretrained weights), which is what the reference's
    constructor does unconditionally.  ``use_bilinear_sampling=True`` keeps the reference's ``grid_sample`` branch (:286-301).
    """

    def __init__(self, layer=['rel1_1', 'relu2_1', 'relu3_1', 'relu4_1'], vgg=None):
        super(PerceptualCorrectness, self).__init__()
        if vgg is None:
            from model.networks.external_function import VGG19      # the reference's, through compat.install()
            vgg = VGG19()
        self.add_module('vgg', vgg)
        self.layer = layer
        self.eps = 1e-8
        from .resample2d import Resample2dCosine
        self.resample_cosine = Resample2dCosine(4, 1, sigma=2, eps=1e-8)

    def __call__(self, target, source, flow_list, used_layers, mask=None, use_bilinear_sampling=False):
        used_layers = sorted(used_layers, reverse=True)
        self.target_vgg, self.source_vgg = self.vgg(target), self.vgg(source)
        loss = 0
        for i in range(len(flow_list)):
            loss += self.calculate_loss(flow_list[i], self.layer[used_layers[i]], mask, use_bilinear_sampling)
        return loss

    def calculate_loss(self, flow, layer, mask=None, use_bilinear_sampling=False):
        tgt, src = self.target_vgg[layer], self.source_vgg[layer]
        b, c, h, w = tgt.shape
        n = h * w
        flow = F.interpolate(flow, [h, w])
        # the best similarity any source position offers each target position (:259-271): N x N correlation of unit vectors
        t_flat = tgt.reshape(b, c, n)
        s_unit = src.reshape(b, c, n).transpose(1, 2)
        s_unit = s_unit / (s_unit.norm(dim=2, keepdim=True) + self.eps)
        t_unit = t_flat / (t_flat.norm(dim=1, keepdim=True) + self.eps)
        best = torch.bmm(s_unit, t_unit).amax(dim=1)                                   # [b, N]
        # the similarity the flow actually achieves (:273-279)
        if use_bilinear_sampling:
            achieved = F.cosine_similarity(self.bilinear_warp(src, flow), t_flat)
        else:
            achieved = self.resample_cosine(src, flow, tgt).reshape(b, n)              # fused: no warped feature tensor
        loss_map = torch.exp(-achieved / (best + self.eps))
        floor = math.exp(-1.0)                                                         # value of a perfect sample
        if mask is None:
            return loss_map.mean() - floor
        m = F.interpolate(mask, size=(h, w)).reshape(-1, n)
        return (m * (loss_map - floor)).sum() / (m.sum() + self.eps)

    def bilinear_warp(self, source, flow):
        """the ``grid_sample`` alternative of the reference (:309-320): pixel offsets -> offsets on the [-1, 1] sampling grid"""
        b, c, h, w = source.shape
        ys, xs = torch.meshgrid(torch.linspace(-1.0, 1.0, h, device=source.device), torch.linspace(-1.0, 1.0, w, device=source.device),
                                indexing="ij")
        base = torch.stack((xs, ys), dim=-1).unsqueeze(0)                              # [1, h, w, (x, y)]
        step = torch.tensor([2.0 / w, 2.0 / h], device=source.device).view(1, 1, 1, 2)
        grid = (base + flow.permute(0, 2, 3, 1).float() * step).type_as(source)
        return F.grid_sample(source, grid).view(b, c, -1)
