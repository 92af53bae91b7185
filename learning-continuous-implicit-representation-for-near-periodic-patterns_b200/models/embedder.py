"""get_embedder with the reference's signature (models/embedder.py:60-90).

Encoding arithmetic lives in csrc/simt_kernels.cuh (npp_base_feature / npp_encode_kernel); the objects returned
here only carry the constants.  In the default 'coords' mode ``embed`` does not build the 1386-wide table at all:
proposal 0 returns the raw [N,2] coordinates, the other proposals an [N,0] tensor and the Fourier embedder is the
identity, so what flows through the unchanged train scripts (cat / reshape / fancy indexing,
NPP_completion/train.py:93-105,164-181) is the coordinate itself and NPP_Net.forward encodes inside the kernels.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _registry
from ._core import EncoderSpec, Plan


class FourierEmbedder:
    """Embedder(input_dims=1, sampling='gaussian') of the reference (embedder.py:6-56)."""

    def __init__(self, multires, res):
        # same call as embedder.py:26 so the default-device RNG stream stays aligned with the reference
        self.freq_bands = torch.normal(mean=0.0, std=1.0, size=(multires, 1)) * 10
        self.freqs = self.freq_bands.detach().reshape(-1).cpu().numpy().astype(np.float32)
        self.out_dim = 1 + 2 * multires
        self.res = res
        self.is_search = False

    def embed(self, inputs):
        if _registry.mode() == "coords":
            return inputs
        outs = [inputs]
        for f in self.freq_bands.to(inputs.device):
            outs += [torch.sin(inputs * f), torch.cos(inputs * f)]
        return torch.cat(outs, -1)


class SearchPositionalEmbedder(FourierEmbedder):
    """Embedder(input_dims=2, is_search=True) of the reference (embedder.py:51-56,76-80): [N,2] coordinates,
    normalised in place, then [u, sin(u f_0), cos(u f_0), ...] -> 2 + 4*multires columns."""

    def __init__(self, multires, res):
        super().__init__(multires, res)
        self.out_dim = 2 + 4 * multires
        self.is_search = True

    def embed(self, inputs):
        if _registry.mode() == "coords":
            return inputs          # NPP_Net_light encodes inside the kernels (npp_encode_search_kernel)
        inputs[:, 0] = ((inputs[:, 0] / self.res[0]) - 0.5) * 2
        inputs[:, 1] = ((inputs[:, 1] / self.res[1]) - 0.5) * 2
        return super().embed(inputs)


class PeriodicEmbedder:
    """Embedder_periodic of the reference (embedder.py:94-148): constants only.  include_input is False in search
    mode (embedder.py:93-95)."""

    def __init__(self, res, selected_angles, selected_periods, freq_scales, freq_offsets, angle_offsets,
                 include_input=True):
        self.res = tuple(int(r) for r in res)
        self.cos_t, self.sin_t, self.period = EncoderSpec.proposal_tables(
            selected_angles, selected_periods, freq_scales, freq_offsets, angle_offsets)
        self.include_input = bool(include_input)
        self.out_dim = 2 * (int(self.include_input) + 2 * self.cos_t.shape[1])
        self.index = _registry.add_periodic(self)
        self._plan = None

    def embed(self, inputs):
        if _registry.mode() == "coords":
            if self.index == 0:
                return inputs[:, :2]
            return inputs.new_zeros((inputs.shape[0], 0))
        # table mode: the 22 base features of this proposal, computed by the CUDA encoder (n_freq = 0)
        if self._plan is None:
            spec = EncoderSpec(res=self.res, cos_t=self.cos_t[None], sin_t=self.sin_t[None], period=self.period[None],
                               freqs=np.zeros((0,), np.float32), include_input=self.include_input)
            self._plan = Plan(spec, max_rows=128, training=False)
        return self._plan.encode(inputs)


def get_embedder(multires, i=0, res=None, selected_angles=None, selected_periods=None,
                 freq_scales=None, freq_offsets=None, angle_offsets=None, is_search=False):
    if i == -1:
        return nn.Identity(), 3
    if selected_periods is None and selected_angles is None:
        e = SearchPositionalEmbedder(multires, res) if is_search else FourierEmbedder(multires, res)
        _registry.begin_session(e)
        return e, e.out_dim
    e = PeriodicEmbedder(res, selected_angles, selected_periods, freq_scales, freq_offsets, angle_offsets,
                         include_input=not is_search)
    return e, e.out_dim


def current_encoder_spec() -> EncoderSpec:
    """EncoderSpec of the embedders created since the last Fourier get_embedder call."""
    if _registry.nerf is None or not _registry.periodic:
        raise RuntimeError("NPP_Net needs get_embedder(...) to be called first for the Fourier embedder and for every "
                           "proposal (as create_npp_net does, reference models/helpers.py:87,108-116)")
    per = _registry.periodic
    return EncoderSpec(res=per[0].res, cos_t=np.stack([p.cos_t for p in per]), sin_t=np.stack([p.sin_t for p in per]),
                       period=np.stack([p.period for p in per]), freqs=_registry.nerf.freqs,
                       include_input=per[0].include_input)
