"""NPP_Net / NPP_Net_top1 with the reference constructor and forward signatures (models/networks.py:8-173),
executed by libnpp_b200 (tcgen05 GEMMs with fused bias / snake epilogues, recomputation-free backward, see
DESIGN.md), and NPP_Net_light (:176-263), the search-stage network, on the same kernels.

Parameters are ``nn.Parameter`` views into one flat fp32 arena owned by the plan, under the reference's state_dict
names, so ``model.parameters()``, ``state_dict()`` / ``load_state_dict()`` and foreign optimisers keep working.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (re-exported like the reference module)

from .activations import *  # noqa: F401,F403
from . import embedder as _embedder
from ._core import Plan, native, plan as _planmod


class _NppFunction(torch.autograd.Function):
    """coords (or a materialised encoding) -> logits; backward fills the plan's gradient arena and publishes the
    per-tensor views of it as ``.grad``.  Activations live in the plan's workspace, so one backward per forward.

    Only a one-element anchor tensor goes through autograd (it makes the engine call ``backward``): handing all 28
    parameters to the Function and 28 gradient views back cost ~0.4 ms of host time per step in AccumulateGrad
    nodes and view construction -- as much as the GPU work of a whole step."""

    @staticmethod
    def forward(ctx, x, net, anchor):
        plan = net._plan_for(x.shape[0])
        net._sync_if_dirty(net._params)
        if x.shape[1] == 2:
            logits = plan.forward(x)
        else:
            logits = plan.forward_encoded(x)
        net._generation += 1
        ctx.net, ctx.n, ctx.generation = net, x.shape[0], net._generation
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        net = ctx.net
        if ctx.generation != net._generation:
            raise RuntimeError("NPP_Net: the activations of this forward were overwritten by a later forward; "
                               "call backward() before running the network again")
        # plan.backward OVERWRITES the gradient arena.  If the arena views are still installed as .grad (a previous
        # backward without zero_grad() in between: micro-batch accumulation, or a foreign training loop), keep what
        # they hold and add it back, like autograd's AccumulateGrad would.
        pending = any(p.grad is g and g is not None for p, g in zip(net._params, net._grad_view_list))
        keep = net._plan.grads[: net._plan.trained_floats].clone() if pending else None
        net._plan.backward(ctx.n, grad_logits)
        if keep is not None:
            net._plan.grads[: net._plan.trained_floats].add_(keep)
        net._publish_grads()
        return None, None, None


class _ArenaLinear(nn.Module):
    """Parameter holder with nn.Linear's attribute names; the tensors are views into the plan arena."""

    def __init__(self, weight, bias):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)
        self.in_features, self.out_features = weight.shape[1], weight.shape[0]

    def forward(self, x):  # only used for the unused alpha_linear / debugging
        return F.linear(x, self.weight, self.bias)


class _FusedNet(nn.Module):
    def _build(self, topk_model, D, W, skips, activation, output_ch, reference_order, light=False):
        if output_ch != 3 or len(skips) != 1:
            raise NotImplementedError("output_ch=3 and a single skip connection are required")
        self.D, self.W, self.skips = D, W, skips
        spec = _embedder.current_encoder_spec()
        if (spec.topk > 1) != topk_model:
            raise ValueError("number of registered proposals does not match the network class")
        if light != (not spec.include_input):
            raise ValueError("NPP_Net_light goes with the search-mode embedders (get_embedder(..., is_search=True)) "
                             "and the other networks with the regular ones")
        self._spec = spec
        # `i in skips` is tested for i < D (forward) and i < D-1 (constructor): a skip index beyond the trunk is inert
        # (the search default D=4 with skips=[4], options/arg_config.py:114).  skips[0] == D-1 is the one value the
        # reference cannot run: forward concatenates after the last trunk layer and feature_linear1 then fails on the
        # shape (networks.py:70-73) -- refuse it instead of training a different network.
        if skips[0] == D - 1:
            raise ValueError(f"skips=[{skips[0]}] with D={D}: the reference forward concatenates the input after the last "
                             "trunk layer and fails in feature_linear1; use a skip index < D-1, or >= D for none")
        self._skip_layer = skips[0] if skips[0] < D - 1 else -1
        self._model_kind = _planmod.MODEL_LIGHT if light else None
        self._activation = 'snake' if activation == 'snake' else 'relu'     # networks.py:51-54: anything else is relu
        self._plan = Plan(spec, depth=D, width=W, skip_layer=self._skip_layer, max_rows=1 << 15,
                          model=self._model_kind, activation=self._activation)
        self._generation = 0
        views = self._plan.param_views()
        # Initialise exactly like the reference: construct nn.Linear modules in the reference's order
        # (networks.py:42-49 / 127-138) so a given torch seed yields the same weights, then copy.
        for name, (out_f, in_f) in reference_order:
            lin = nn.Linear(in_f, out_f)
            views[name + ".weight"].copy_(lin.weight.detach())
            views[name + ".bias"].copy_(lin.bias.detach())
        self._plan.sync_weights()

        def holder(name):
            return _ArenaLinear(views[name + ".weight"], views[name + ".bias"])

        self.periodic_linears = nn.ModuleList([holder(f"periodic_linears.{i}") for i in range(D)])
        if topk_model or light:
            self.scale_linears = nn.ModuleList([holder("scale_linears.0")])
        self.pos_linears = nn.ModuleList([holder("pos_linears.0")])
        self.feature_linear1 = holder("feature_linear1")
        self.feature_linear2 = holder("feature_linear2")
        self.alpha_linear = holder("alpha_linear")
        self.rgb_linear = holder("rgb_linear")
        self.snakes = SnakeActivation() if self._activation == 'snake' else None
        named = dict(self.named_parameters())
        self._param_names = list(named.keys())
        self._params = [named[k] for k in self._param_names]
        self._versions = None
        gv = self._plan.grad_views()          # persistent views of the gradient arena, one per trained tensor
        self._grad_view_list = [gv.get(k) for k in self._param_names]
        # not a Parameter (state_dict / parameters() stay the reference's): see _NppFunction
        object.__setattr__(self, "_anchor", torch.zeros(1, device=self._plan.device, requires_grad=True))

    # --------------------------------------------------------------------------------------------
    def _plan_for(self, n):
        if n > self._plan.max_rows:
            old = self._plan
            cap = 1 << (int(n) - 1).bit_length()
            self._plan = Plan(self._spec, depth=self.D, width=self.W, skip_layer=self._skip_layer, max_rows=cap,
                              arenas=(old.params, old.grads, old.exp_avg, old.exp_avg_sq), model=self._model_kind,
                              activation=self._activation)
            self._plan.adam_steps = old.adam_steps
            self._plan.sync_weights()
            old.close()
        return self._plan

    def _sync_if_dirty(self, params):
        """Refresh the fp16 shadow weights if anyone but the fused optimiser wrote the fp32 parameters."""
        versions = tuple(p._version for p in params)
        if versions != self._versions:
            self._plan.sync_weights()
            self._versions = versions

    def _publish_grads(self):
        """After plan.backward: make ``p.grad`` the arena view of every trained tensor.  A gradient somebody else put
        there in the meantime (a second loss through other modules) is added to, like autograd would."""
        for p, g in zip(self._params, self._grad_view_list):
            if g is None or not p.requires_grad:
                continue
            if p.grad is None or p.grad is g:
                p.grad = g
            else:
                p.grad = p.grad + g

    def mark_clean(self):
        self._versions = tuple(p._version for p in self._params)

    def _apply(self, fn, *a, **k):
        before = self._plan.params.data_ptr()
        out = super()._apply(fn, *a, **k)
        if any(p.device != self._plan.device or p.dtype != torch.float32 for p in self._params) or \
                self._plan.params.data_ptr() != before:
            raise RuntimeError("NPP_Net parameters live in a CUDA fp32 arena and cannot be moved or cast")
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._plan.sync_weights()
        self.mark_clean()
        return out

    @torch.no_grad()
    def render_into(self, coords, image, normalize_type=1):
        """Inference straight into an image (the evaluation loop of NPP_completion/train.py:277-309 in one call):
        image[..., y, x, :] = sigmoid / tanh of the network output for every (row, col) in `coords` [N,2]."""
        plan = self._plan_for(min(int(coords.shape[0]), self._plan.max_rows))
        self._sync_if_dirty(self._params)
        self._generation += 1          # the activations of a pending forward are overwritten
        return plan.render_into(coords, image, normalize_type)

    def forward(self, x, x_periodic):
        """x is ignored (None) exactly like the reference; x_periodic is either [N,2] coordinates ('coords' embed
        mode) or the materialised [N, K*462] encoding."""
        if x_periodic.shape[1] not in (2, self._plan.encoding_width):
            raise AssertionError(f"x_periodic has {x_periodic.shape[1]} columns; expected 2 (coordinates) or "
                                 f"{self._plan.encoding_width} (materialised encoding)")
        if x_periodic.shape[0] == 0:
            return x_periodic.new_zeros((0, 3))
        if x_periodic.dtype != torch.float32:
            x_periodic = x_periodic.float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params):
            return _NppFunction.apply(x_periodic, self, self._anchor)
        return _NppFunction.apply(x_periodic, self, None)


class NPP_Net(_FusedNet):
    def __init__(self, input_ch_periodic, input_ch_periodic_aux, freq_scales, freq_offsets, angle_offsets, D=8, W=256,
                 freq_nerf=3, output_ch=3, skips=[4], activation='relu'):
        super().__init__()
        self.scale, self.offset, self.angle_offset = len(freq_scales), len(freq_offsets), len(angle_offsets)
        ch = int(input_ch_periodic) * freq_nerf
        aux = int(input_ch_periodic_aux) * freq_nerf
        self.input_ch_periodic, self.input_ch_periodic_aux = ch, aux
        order = [(f"periodic_linears.{i}", (W, ch if i == 0 else (W + ch if (i - 1) in skips else W))) for i in range(D)]
        order += [("scale_linears.0", (W, aux + W)), ("pos_linears.0", (W // 2, 2 * W)),
                  ("feature_linear1", (W, W)), ("feature_linear2", (W, W)), ("alpha_linear", (1, W)),
                  ("rgb_linear", (output_ch, W // 2))]
        self._build(True, D, W, list(skips), activation, output_ch, order)
        assert self._plan.encoding_width == ch + aux, "embedder widths do not match input_ch_periodic(_aux)"


class NPP_Net_top1(_FusedNet):
    def __init__(self, input_ch_periodic, freq_scales, freq_offsets, angle_offsets, D=8, W=256, freq_nerf=3,
                 output_ch=3, skips=[4], activation='relu'):
        super().__init__()
        self.scale, self.offset, self.angle_offset = len(freq_scales), len(freq_offsets), len(angle_offsets)
        ch = int(input_ch_periodic) * freq_nerf
        self.input_ch_periodic = ch
        order = [(f"periodic_linears.{i}", (W, ch if i == 0 else (W + ch if (i - 1) in skips else W))) for i in range(D)]
        order += [("pos_linears.0", (W // 2, W)), ("feature_linear1", (W, W)), ("feature_linear2", (W, W)),
                  ("alpha_linear", (1, W)), ("rgb_linear", (output_ch, W // 2))]
        self._build(False, D, W, list(skips), activation, output_ch, order)
        assert self._plan.encoding_width == ch


class NPP_Net_light(_FusedNet):
    """The search-stage network (networks.py:176-263) for len(freq_scales) == 1, where the scale MLP is skipped:
    periodic trunk -> feature_linear1 -> cat(feature1, x) -> pos_linears.0 -> rgb_linear.  `x` is the 2-D positional
    encoding and `x_periodic` the periodic one; in 'coords' embed mode both are the raw [N,2] coordinates."""

    def __init__(self, input_ch_periodic, freq_scales, freq_offsets, angle_offsets, D=8, W=256, input_ch=3,
                 output_ch=3, skips=[4], activation='relu'):
        super().__init__()
        self.scale, self.offset, self.angle_offset = len(freq_scales), len(freq_offsets), len(angle_offsets)
        if self.scale != 1:
            raise NotImplementedError("NPP_Net_light is built for len(freq_scales) == 1 (the reference default, "
                                      "options/arg_config.py:18), where the scale MLP is not executed")
        ch = 2 * (2 * self.offset * self.angle_offset)            # networks.py:188
        if int(input_ch_periodic) != ch:
            raise ValueError(f"input_ch_periodic={input_ch_periodic} but the search-mode periodic encoding has {ch} columns")
        self.input_ch, self.input_ch_periodic = int(input_ch), ch
        order = [(f"periodic_linears.{i}", (W, ch if i == 0 else (W + ch if (i - 1) in skips else W))) for i in range(D)]
        order += [("scale_linears.0", (W, W)), ("pos_linears.0", (W // 2, self.input_ch + W)),
                  ("feature_linear1", (W, W)), ("feature_linear2", (W, W)), ("alpha_linear", (1, W)),
                  ("rgb_linear", (output_ch, W // 2))]
        self._build(False, D, W, list(skips), activation, output_ch, order, light=True)
        assert self._plan.encoding_width == ch + self.input_ch, "embedder widths do not match input_ch(_periodic)"

    def forward(self, x, x_periodic):
        if x_periodic.shape[1] == 2:
            return super().forward(None, x_periodic)
        # materialised encodings: the kernels take [periodic | positional] (npp_forward_encoded)
        return super().forward(None, torch.cat([x_periodic.float(), x.float().to(x_periodic.device)], -1))
