"""B200-native NPP-Net training hot path (encoding -> MLP forward/backward -> loss -> Adam).

See DESIGN.md for the scope and include/npp_b200.h for the C ABI this package wraps.
"""
from . import _native  # noqa: F401

__all__ = ["_native"]
