"""Activation modules the reference exports from models/activations.py.  Only SnakeActivation (a=1) is on the
hot path, where it is fused into the GEMM epilogue (csrc/gemm_sm100.cuh, EPI_SNAKE); the module here exists so
that ``from models.activations import *`` keeps providing the same names."""
import torch
import torch.nn as nn


class SinActivation(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class SnakeActivation(nn.Module):
    """x + sin(a x)^2 / a  (reference models/activations.py:29-35)."""

    def __init__(self, a=1):
        super().__init__()
        self.a = a

    def forward(self, x):
        return x + torch.square(torch.sin(self.a * x)) / self.a
