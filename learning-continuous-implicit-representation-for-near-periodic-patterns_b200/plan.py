"""Python owner of one ``NppPlan`` (include/npp_b200.h): encoder constants, the fp32 parameter /
gradient / Adam arenas (torch tensors, so ``nn.Parameter`` views and ``torch.distributed`` work on
them) and thin methods over the C ABI.  All compute happens in libnpp_b200.so."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as nat

MODEL_TOPK = 0
MODEL_TOP1 = 1
MODEL_LIGHT = 2      # NPP_Net_light with the search-mode encoders (include/npp_b200.h)


@dataclass
class EncoderSpec:
    """Constants of the periodicity-aware encoding, derived with the reference's own torch fp32 ops
    (models/embedder.py:112-127) so that cos/sin(theta) and the periods match it bit for bit."""
    res: Sequence[int]                       # (H, W)
    cos_t: np.ndarray                        # [K, 2, n_aug] float32
    sin_t: np.ndarray
    period: np.ndarray
    freqs: np.ndarray                        # [n_freq] float32, Gaussian Fourier frequencies
    include_input: bool = True

    @property
    def topk(self) -> int:
        return int(self.cos_t.shape[0])

    @property
    def n_aug(self) -> int:
        return int(self.cos_t.shape[2])

    @property
    def base_width(self) -> int:             # 22 at the defaults
        return 2 * (int(self.include_input) + 2 * self.n_aug)

    @property
    def width_per_proposal(self) -> int:     # 462 at the defaults
        return self.base_width * (1 + 2 * len(self.freqs))

    @staticmethod
    def proposal_tables(selected_angles, selected_periods, freq_scales, freq_offsets, angle_offsets):
        """One proposal -> (cos_t, sin_t, period) each [2, n_aug]; loop order of embedder.py:117-120."""
        angles = torch.as_tensor(selected_angles, dtype=torch.float32).cpu()
        freq_bands = torch.as_tensor(selected_periods, dtype=torch.float32).cpu()
        n_aug = len(freq_scales) * len(freq_offsets) * len(angle_offsets)
        cos_t = np.zeros((2, n_aug), np.float32)
        sin_t = np.zeros((2, n_aug), np.float32)
        period = np.zeros((2, n_aug), np.float32)
        for idx in range(2):
            a = 0
            for fs in freq_scales:
                for fo in freq_offsets:
                    for ao in angle_offsets:
                        freq = (freq_bands[idx] + fo) * fs
                        theta = torch.deg2rad(angles[idx] + ao)
                        cos_t[idx, a] = torch.cos(theta).item()
                        sin_t[idx, a] = torch.sin(theta).item()
                        period[idx, a] = freq.item()
                        a += 1
        return cos_t, sin_t, period

    @classmethod
    def from_proposals(cls, res, selected_angles, selected_periods, freqs, freq_scales=(1,),
                       freq_offsets=(0, -1, 1, 0.5, -0.5), angle_offsets=(0,), include_input=True):
        tabs = [cls.proposal_tables(a, p, freq_scales, freq_offsets, angle_offsets)
                for a, p in zip(selected_angles, selected_periods)]
        return cls(res=tuple(int(r) for r in res),
                   cos_t=np.stack([t[0] for t in tabs]), sin_t=np.stack([t[1] for t in tabs]),
                   period=np.stack([t[2] for t in tabs]),
                   freqs=np.asarray(freqs, np.float32).reshape(-1), include_input=include_input)


@dataclass
class TensorSlot:
    name: str
    offset: int
    shape: tuple
    trained: bool


class Plan:
    """One fused NPP-Net instance on the current CUDA device."""

    def __init__(self, enc: EncoderSpec, *, depth: int = 8, width: int = 512, skip_layer: int = 4,
                 max_rows: int = 1 << 16, wgrad_splits: int = 0, training: bool = True, arenas=None,
                 model: Optional[int] = None, activation: str = "snake"):
        if not torch.cuda.is_available():
            raise nat.NppError("npp_b200 needs a CUDA device (B200, sm_100); there is no CPU fallback")
        self.lib = nat.lib()
        self.enc = enc
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_rows = int(max_rows)
        self.topk = enc.topk
        self.model = (MODEL_TOPK if enc.topk > 1 else MODEL_TOP1) if model is None else int(model)
        cfg = nat.NppConfig()
        cfg.model = self.model
        cfg.topk = enc.topk
        cfg.depth = depth
        cfg.width = width
        cfg.skip_layer = skip_layer
        cfg.n_aug = enc.n_aug
        cfg.n_freq = len(enc.freqs)
        cfg.include_input = int(enc.include_input)
        cfg.res_h, cfg.res_w = int(enc.res[0]), int(enc.res[1])
        cfg.wgrad_splits = wgrad_splits
        cfg.activation = 0 if activation == "snake" else 1      # anything else is relu (models/networks.py:51-54)
        self.activation = "snake" if activation == "snake" else "relu"
        cfg.max_rows = self.max_rows
        keep = [np.ascontiguousarray(a, np.float32) for a in (enc.cos_t, enc.sin_t, enc.period, enc.freqs)]
        fp = C.POINTER(C.c_float)
        cfg.cos_t, cfg.sin_t, cfg.period, cfg.freq = (a.ctypes.data_as(fp) for a in keep)
        handle = C.c_void_p()
        nat.check(self.lib.npp_plan_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        total, trained = C.c_int64(), C.c_int64()
        nat.check(self.lib.npp_plan_arena_floats(handle, C.byref(total), C.byref(trained)))
        self.arena_floats, self.trained_floats = total.value, trained.value
        self.slots: List[TensorSlot] = []
        info = nat.NppTensorInfo()
        for i in range(self.lib.npp_plan_tensor_count(handle)):
            nat.check(self.lib.npp_plan_tensor_info(handle, i, C.byref(info)))
            shape = (info.cols,) if info.is_bias else (info.rows, info.cols)
            self.slots.append(TensorSlot(info.name.decode(), info.offset, shape, bool(info.trained)))
        if arenas is not None:      # re-plan with a larger workspace around existing parameters / Adam state
            self.params, self.grads, self.exp_avg, self.exp_avg_sq = arenas
            assert self.params.numel() == self.arena_floats
        else:
            self.params = torch.zeros(self.arena_floats, device=self.device)
            self.grads = torch.zeros(self.arena_floats, device=self.device) if training else None
            self.exp_avg = torch.zeros(self.arena_floats, device=self.device) if training else None
            self.exp_avg_sq = torch.zeros(self.arena_floats, device=self.device) if training else None
        nat.check(self.lib.npp_plan_bind(handle, nat.ptr(self.params), nat.ptr(self.grads),
                                         nat.ptr(self.exp_avg), nat.ptr(self.exp_avg_sq)))
        self.encoding_width = self.lib.npp_plan_encoding_width(handle)
        # dense layers in execution order; layer i owns the debug buffers h{i}, d{i}, delta{i}
        self.layer_names = [s.name[:-len(".weight")] for s in self.slots
                            if s.trained and s.name.endswith(".weight") and not s.name.startswith("rgb_linear")]
        self.adam_steps = 0

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "handle", None):
            self.lib.npp_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- parameters
    def view(self, arena: torch.Tensor, slot: TensorSlot) -> torch.Tensor:
        n = int(np.prod(slot.shape))
        return arena[slot.offset: slot.offset + n].view(*slot.shape)

    def param_views(self) -> Dict[str, torch.Tensor]:
        return {s.name: self.view(self.params, s) for s in self.slots}

    def grad_views(self) -> Dict[str, torch.Tensor]:
        return {s.name: self.view(self.grads, s) for s in self.slots if s.trained}

    def load_state(self, state: Dict[str, "np.ndarray | torch.Tensor"], strict: bool = True):
        """Copy reference-shaped tensors (state_dict keys of NPP_Net / NPP_Net_top1) into the arena."""
        views = self.param_views()
        for name, v in views.items():
            if name not in state:
                if strict:
                    raise KeyError(name)
                continue
            src = torch.as_tensor(np.asarray(state[name]) if not torch.is_tensor(state[name]) else state[name])
            v.copy_(src.to(self.device, torch.float32).reshape(v.shape))
        self.sync_weights()

    def reset_parameters(self, seed: Optional[int] = None, on_device: bool = False):
        """torch.nn.Linear default init for every tensor: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias
        (the reference keeps PyTorch's default because weights_init_normal only touches Conv/BatchNorm,
        models/helpers.py:65-71,140-141).  on_device: draw on the GPU (a fresh fit without a host round trip; another
        random stream than the CPU generator) and clear the Adam state."""
        gen = torch.Generator(device=self.device if on_device else "cpu")
        if seed is not None:
            gen.manual_seed(seed)
        views = self.param_views()
        by_name = dict((t.name, t) for t in self.slots)
        for s in self.slots:
            fan_in = s.shape[1] if len(s.shape) == 2 else by_name[s.name.replace(".bias", ".weight")].shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            if on_device:
                views[s.name].copy_((torch.rand(s.shape, generator=gen, device=self.device) * 2 - 1) * bound)
            else:
                views[s.name].copy_((torch.rand(s.shape, generator=gen) * 2 - 1) * bound)
        if on_device and self.exp_avg is not None:
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
        self.sync_weights()

    def state(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().clone() for k, v in self.param_views().items()}

    def sync_weights(self):
        nat.check(self.lib.npp_sync_weights(self.handle, nat.current_stream()))

    # ------------------------------------------------------------------- compute
    def _coords(self, coords: torch.Tensor) -> torch.Tensor:
        c = coords.to(self.device, torch.float32).contiguous()
        if c.dim() != 2 or c.shape[1] != 2:
            raise ValueError("coords must be [N,2] (row, col)")
        if c.shape[0] > self.max_rows:
            raise ValueError(f"{c.shape[0]} rows exceed the plan capacity max_rows={self.max_rows}")
        return c

    def encode(self, coords: torch.Tensor) -> torch.Tensor:
        c = coords.to(self.device, torch.float32).contiguous()      # no workspace involved: any row count
        out = torch.empty(c.shape[0], self.encoding_width, device=self.device)
        if c.shape[0]:
            nat.check(self.lib.npp_encode(self.handle, c.data_ptr(), c.shape[0], out.data_ptr(), nat.current_stream()))
        return out

    def forward(self, coords: torch.Tensor) -> torch.Tensor:
        c = self._coords(coords)
        logits = torch.empty(c.shape[0], 3, device=self.device)
        if c.shape[0]:
            nat.check(self.lib.npp_forward(self.handle, c.data_ptr(), c.shape[0], logits.data_ptr(), nat.current_stream()))
        return logits

    def render_into(self, coords: torch.Tensor, image: torch.Tensor, normalize_type: int = 1) -> torch.Tensor:
        """Inference straight into an image: image[y, x, :] = sigmoid(net(y, x)) for every coordinate row (any number
        of rows).  image: contiguous fp32 CUDA [H, W, 3] (or [1, H, W, 3] as the train scripts hold it)."""
        c = coords.to(self.device, torch.float32).contiguous()
        if c.dim() != 2 or c.shape[1] != 2:
            raise ValueError("coords must be [N,2] (row, col)")
        if not (image.is_cuda and image.dtype == torch.float32 and image.is_contiguous() and image.shape[-1] == 3
                and image.dim() in (3, 4) and (image.dim() == 3 or image.shape[0] == 1)):
            raise ValueError("image must be a contiguous fp32 CUDA tensor [H, W, 3] or [1, H, W, 3]")
        h, w = int(image.shape[-3]), int(image.shape[-2])
        if c.shape[0]:
            nat.check(self.lib.npp_render_into(self.handle, c.data_ptr(), c.shape[0], image.data_ptr(), h, w,
                                               int(normalize_type), nat.current_stream()))
        return image

    def forward_encoded(self, enc: torch.Tensor) -> torch.Tensor:
        """Forward on a materialised [N, K*462] fp32 encoding (reference layout)."""
        e = enc.to(self.device, torch.float32).contiguous()
        if e.dim() != 2 or e.shape[1] != self.encoding_width:
            raise ValueError(f"encoding must be [N,{self.encoding_width}], got {tuple(e.shape)}")
        if e.shape[0] > self.max_rows:
            raise ValueError(f"{e.shape[0]} rows exceed the plan capacity max_rows={self.max_rows}")
        logits = torch.empty(e.shape[0], 3, device=self.device)
        if e.shape[0]:
            nat.check(self.lib.npp_forward_encoded(self.handle, e.data_ptr(), e.shape[0], logits.data_ptr(),
                                                   nat.current_stream()))
        return logits

    def backward(self, n: int, grad_logits: torch.Tensor):
        g = grad_logits.to(self.device, torch.float32).contiguous()
        assert g.shape == (n, 3)
        nat.check(self.lib.npp_backward(self.handle, n, g.data_ptr(), nat.current_stream()))

    def mse(self, logits, target, mask=None, n_norm: Optional[int] = None, want_pred: bool = False):
        n = logits.shape[0]
        n_norm = n if n_norm is None else int(n_norm)
        g = torch.empty_like(logits)
        loss = torch.zeros((), device=self.device)
        pred = torch.empty_like(logits) if want_pred else None
        nat.check(self.lib.npp_mse_fwd_bwd(self.handle, logits.data_ptr(), target.data_ptr(), nat.ptr(mask), n, n_norm,
                                           nat.ptr(pred), g.data_ptr(), loss.data_ptr(), nat.current_stream()))
        return loss, g, pred

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, step: Optional[int] = None):
        if step is None:
            self.adam_steps += 1
            step = self.adam_steps
        nat.check(self.lib.npp_adam_step(self.handle, lr, betas[0], betas[1], eps, step, nat.current_stream()))

    def train_step(self, coords, target, mask, lr, loss_out: torch.Tensor, n_norm: Optional[int] = None,
                   betas=(0.9, 0.999), eps: float = 1e-8, step: Optional[int] = None):
        """encode + forward + sigmoid/masked-MSE + backward + Adam, all inside libnpp_b200.
        Inputs must already be contiguous fp32 CUDA tensors (no host work on this path)."""
        if step is None:
            self.adam_steps += 1
            step = self.adam_steps
        n = coords.shape[0]
        nat.check(self.lib.npp_train_step(self.handle, coords.data_ptr(), target.data_ptr(), nat.ptr(mask), n,
                                          n if n_norm is None else int(n_norm), lr, betas[0], betas[1], eps, step,
                                          loss_out.data_ptr(), nat.current_stream()))

    # ---- the step in three phases (data parallelism: the gradient arena is summed over the ranks in between)
    def step_forward_backward(self, coords, target, mask, n_norm: int):
        nat.check(self.lib.npp_step_forward_backward(self.handle, coords.data_ptr(), target.data_ptr(), nat.ptr(mask),
                                                     coords.shape[0], int(n_norm), nat.current_stream()))

    def layer_count(self) -> int:
        return self.lib.npp_plan_layer_count(self.handle)

    def grad_range(self, layer_begin: int, layer_end: int) -> torch.Tensor:
        """View of the gradient arena holding weights + biases of the dense layers [layer_begin, layer_end) in
        execution order (rgb_linear rides with the last layer)."""
        off, cnt = C.c_int64(), C.c_int64()
        nat.check(self.lib.npp_plan_layer_grad_range(self.handle, layer_begin, layer_end, C.byref(off), C.byref(cnt)))
        return self.grads[off.value: off.value + cnt.value]

    def step_wgrad(self, layer_begin: int, layer_end: int, n: int, n_norm: int):
        nat.check(self.lib.npp_step_wgrad(self.handle, layer_begin, layer_end, int(n), int(n_norm), nat.current_stream()))

    def step_finish(self, n_norm: int, lr: float, loss_out: torch.Tensor, betas=(0.9, 0.999), eps: float = 1e-8,
                    step: Optional[int] = None):
        if step is None:
            self.adam_steps += 1
            step = self.adam_steps
        nat.check(self.lib.npp_step_finish(self.handle, int(n_norm), lr, betas[0], betas[1], eps, step,
                                           loss_out.data_ptr(), nat.current_stream()))

    def fit_run(self, coords_all: torch.Tensor, target_all: torch.Tensor, mask_all: Optional[torch.Tensor] = None, *,
                lrate: float = 5e-4, lrate_decay: float = 500, decay_rate: float = 0.1, betas=(0.9, 0.999),
                eps: float = 1e-8, losses: Optional[torch.Tensor] = None, stream: Optional[int] = None) -> torch.Tensor:
        """`iters` train steps in one call (npp_fit_run): coords_all [iters, n, 2], target_all [iters, n, 3], optional
        mask_all [iters, n, 1], contiguous fp32 CUDA tensors.  The learning rate follows the reference loops' rewrite
        (decay_steps = lrate_decay * 100, NPP_proposal/search.py:139-144).  Returns the per-step losses [iters]."""
        iters, n = int(coords_all.shape[0]), int(coords_all.shape[1])
        for t, last in ((coords_all, 2), (target_all, 3)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[:2] == (iters, n)
                    and t.shape[2] == last):
                raise ValueError("fit_run needs contiguous fp32 CUDA tensors [iters, n, 2] and [iters, n, 3]")
        if mask_all is not None and not (mask_all.is_cuda and mask_all.dtype == torch.float32 and
                                         mask_all.is_contiguous() and mask_all.numel() == iters * n):
            raise ValueError("mask_all must be a contiguous fp32 CUDA tensor [iters, n, 1]")
        if n > self.max_rows:
            raise ValueError(f"{n} rows exceed the plan capacity max_rows={self.max_rows}")
        if losses is None:
            losses = torch.zeros(iters, device=self.device)
        first = self.adam_steps + 1
        nat.check(self.lib.npp_fit_run(self.handle, coords_all.data_ptr(), target_all.data_ptr(), nat.ptr(mask_all), n,
                                       iters, lrate, decay_rate, float(lrate_decay) * 100.0, betas[0], betas[1], eps,
                                       first, losses.data_ptr(), nat.current_stream() if stream is None else stream))
        self.adam_steps += iters
        return losses

    def prefetch_encode(self, coords: torch.Tensor):
        """Encode the coordinates of a FUTURE `train_step` on the plan's side stream, so that it overlaps with the
        step that is enqueued next (call this first, then that step).  `coords` must be the very tensor (same storage,
        unchanged) the future step is called with; it must be a contiguous fp32 CUDA tensor."""
        nat.check(self.lib.npp_encode_prefetch(self.handle, coords.data_ptr(), coords.shape[0], nat.current_stream()))

    def keep_grads(self, on: bool = True):
        nat.check(self.lib.npp_set_keep_grads(self.handle, int(on)))

    def launch_count(self) -> int:
        return self.lib.npp_last_launch_count(self.handle)

    PROFILE_CLASSES = ("encode", "gemm_fwd", "head_loss", "gemm_dgrad", "gemm_wgrad", "grad_finalize", "adam_shadow")

    def profile(self, on: bool):
        nat.check(self.lib.npp_profile_enable(self.handle, int(on)))

    def profile_read(self):
        """{class: (milliseconds, launches)} accumulated since profile(True); synchronises."""
        k = len(self.PROFILE_CLASSES)
        ms = (C.c_double * k)()
        ln = (C.c_int64 * k)()
        nat.check(self.lib.npp_profile_read(self.handle, k, ms, ln))
        return {name: (ms[i], ln[i]) for i, name in enumerate(self.PROFILE_CLASSES)}

    # --------------------------------------------------------------------- tests
    def debug(self, name: str, n: int) -> torch.Tensor:
        w = self.lib.npp_debug_width(self.handle, name.encode())
        if w < 0:
            raise KeyError(name)
        out = torch.empty(n, w, device=self.device)
        nat.check(self.lib.npp_debug_copy(self.handle, name.encode(), n, out.data_ptr(), nat.current_stream()))
        return out

    def grad_scale(self) -> float:
        return float(self.lib.npp_debug_grad_scale(self.handle, nat.current_stream()))
