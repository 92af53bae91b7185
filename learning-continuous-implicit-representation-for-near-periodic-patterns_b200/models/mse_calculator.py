"""img2mse with the reference signature (models/mse_calculator.py:13-27).  Like the reference module this file is
star-imported by the train scripts, which rely on it for ``os``, ``torch``, ``np`` and the activation classes."""
import os  # noqa: F401

import numpy as np
import torch

from .activations import *  # noqa: F401,F403
from ._core import NppAdaptiveLoss, fused_l2_img2mse


def img2mse(x, y, loss_type, adaptive, mask=None):
    if loss_type == 'l2' and x.is_cuda and x.dim() == 2 and x.shape[1] == 3 and x.shape == y.shape and \
            y.is_cuda and (mask is None or (mask.is_cuda and mask.numel() == x.shape[0])):
        return fused_l2_img2mse(x, y, mask)                     # one CUDA pass: loss + dL/dx
    diff = x - y
    if mask is not None:
        diff = diff * mask + (1 - mask) * diff * 0.3
    if loss_type == 'l2':
        loss = diff ** 2
    elif loss_type == 'robust_loss':
        import robust_loss_pytorch.general          # vendored by the reference (externel_lib/), stays PyTorch
        loss = torch.mean(robust_loss_pytorch.general.lossfun(
            diff, alpha=torch.Tensor([2.]), scale=torch.Tensor([0.1])))
    elif loss_type == 'robust_loss_adaptive':
        if isinstance(adaptive, NppAdaptiveLoss) and x.is_cuda and x.dim() == 2 and x.shape[1] == 3 and \
                adaptive.num_dims == 3 and (mask is None or mask.numel() == x.shape[0]):
            return adaptive.fused_img2mse(x, y, mask)           # one CUDA pass: loss + every gradient
        loss = torch.mean(adaptive.lossfun(diff))                # any other AdaptiveLossFunction-like object / shape
    else:
        raise ValueError(loss_type)
    return torch.mean(loss)


mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.Tensor([10.]))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)
