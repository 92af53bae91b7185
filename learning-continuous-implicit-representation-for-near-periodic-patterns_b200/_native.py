"""ctypes binding of libnpp_b200.so (C ABI declared in include/npp_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``build_native()`` below) with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no CPU or PyTorch fallback: if the
library is missing or the device is not an sm_100 part, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
CSRC = _HERE / "csrc"
LIB_PATH = CSRC / "libnpp_b200.so"
HEADER = _HERE.parent / "include" / "npp_b200.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


class NppError(RuntimeError):
    pass


class NppConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("topk", C.c_int32), ("depth", C.c_int32), ("width", C.c_int32),
        ("skip_layer", C.c_int32), ("n_aug", C.c_int32), ("n_freq", C.c_int32), ("include_input", C.c_int32),
        ("res_h", C.c_int32), ("res_w", C.c_int32), ("wgrad_splits", C.c_int32), ("activation", C.c_int32),
        ("max_rows", C.c_int64),
        ("cos_t", C.POINTER(C.c_float)), ("sin_t", C.POINTER(C.c_float)),
        ("period", C.POINTER(C.c_float)), ("freq", C.POINTER(C.c_float)),
    ]


class NppTensorInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 64), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32),
        ("is_bias", C.c_int32), ("trained", C.c_int32),
    ]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header.
_P = C.c_void_p
SIGNATURES = {
    "npp_last_error": (C.c_char_p, []),
    "npp_abi_version": (C.c_int, []),
    "npp_plan_create": (C.c_int, [C.POINTER(NppConfig), C.POINTER(_P)]),
    "npp_plan_destroy": (C.c_int, [_P]),
    "npp_plan_arena_floats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "npp_plan_tensor_count": (C.c_int, [_P]),
    "npp_plan_tensor_info": (C.c_int, [_P, C.c_int, C.POINTER(NppTensorInfo)]),
    "npp_plan_encoding_width": (C.c_int, [_P]),
    "npp_plan_bind": (C.c_int, [_P, _P, _P, _P, _P]),
    "npp_sync_weights": (C.c_int, [_P, _P]),
    "npp_encode": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "npp_forward": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "npp_forward_encoded": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "npp_render_into": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "npp_backward": (C.c_int, [_P, C.c_int64, _P, _P]),
    "npp_mse_fwd_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P]),
    "npp_adam_step": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _P]),
    "npp_encode_prefetch": (C.c_int, [_P, _P, C.c_int64, _P]),
    "npp_adam_flat": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _P]),
    "npp_gather_windows": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "npp_sampler_candidates": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int64), C.c_int32,
                                         C.c_int32, C.c_float, _P, _P]),
    "npp_l2_fwd_bwd": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    "npp_robust_adaptive_fwd_bwd": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_int, C.c_float, _P, _P, _P, _P]),
    "npp_train_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_int64, _P, _P]),
    "npp_step_forward_backward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P]),
    "npp_plan_layer_count": (C.c_int, [_P]),
    "npp_plan_layer_grad_range": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "npp_step_wgrad": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P]),
    "npp_step_finish": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _P, _P]),
    "npp_fit_run": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                              C.c_float, C.c_float, C.c_int64, _P, _P]),
    "npp_multi_fit_run": (C.c_int, [C.POINTER(_P), C.c_int32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.c_int64,
                                    C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.POINTER(C.c_int64), C.POINTER(_P), _P]),
    "npp_set_keep_grads": (C.c_int, [_P, C.c_int]),
    "npp_last_launch_count": (C.c_int, [_P]),
    "npp_profile_enable": (C.c_int, [_P, C.c_int]),
    "npp_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "npp_debug_width": (C.c_int, [_P, C.c_char_p]),
    "npp_debug_copy": (C.c_int, [_P, C.c_char_p, C.c_int64, _P, _P]),
    "npp_debug_grad_scale": (C.c_float, [_P, _P]),
    "npp_debug_gemm": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "npp_debug_gemm_bench": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "npp_debug_wgrad": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into csrc/libnpp_b200.so for sm_100a (cross-compiles without a GPU)."""
    srcs = [CSRC / "npp_api.cu"]
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HEADER]
    if not force and LIB_PATH.exists():
        newest = max(p.stat().st_mtime for p in deps)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise NppError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NppError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc).  This package has no CPU / PyTorch fallback path.")
        # NPP_B200_LIB: another build of the same library (kernel A/B experiments); the default is the in-tree build
        handle = C.CDLL(os.environ.get("NPP_B200_LIB") or str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("NPP_B200_LIB") and not hasattr(handle, name):
                continue                # an older experimental build: entry points added since are simply absent
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().npp_last_error()
        raise NppError(msg.decode() if msg else f"libnpp_b200 call failed with code {rc}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
