"""create_npp_net / render / run_network / batchify with the reference signatures (models/helpers.py:14-175).
Star-imported by the train scripts, which take ``torch``, ``np``, ``nn``, ``F``, ``device`` and ``adaptive_pix``
from here."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from .mse_calculator import *  # noqa: F401,F403
from .embedder import *  # noqa: F401,F403
from .networks import *  # noqa: F401,F403
from .networks import NPP_Net, NPP_Net_top1, NPP_Net_light
from .embedder import get_embedder
from .optim import NppAdam

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

# The adaptive robust pixel loss (reference models/helpers.py:8-9: AdaptiveLossFunction(num_dims=3, float32, device=0)):
# same parameters and methods, fused CUDA evaluation inside img2mse.
from ._core import NppAdaptiveLoss as AdaptiveLossFunction  # noqa: E402
adaptive_pix = AdaptiveLossFunction(num_dims=3, float_dtype=np.float32, device=0) if torch.cuda.is_available() else None


def batchify(fn, chunk):
    """Apply fn to row chunks (models/helpers.py:14-25)."""
    if chunk is None:
        return fn

    def ret(inputs, inputs_periodic):
        n = inputs_periodic.shape[0]
        outs = [fn(None if inputs is None else inputs[i:i + chunk], inputs_periodic[i:i + chunk])
                for i in range(0, n, chunk)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
    return ret


def run_network(inputs, inputs_periodic, fn, netchunk=1024 * 64):
    outputs_flat = batchify(fn, netchunk)(inputs, inputs_periodic)
    return torch.reshape(outputs_flat, list(inputs_periodic.shape[:-1]) + [outputs_flat.shape[-1]])


def render(select_coords_emb, select_coords_emb_periodic, args, network_query_fn, network_fn):
    """Network output squashed to RGB (models/helpers.py:41-62)."""
    raw = network_query_fn(select_coords_emb, select_coords_emb_periodic, network_fn)
    if args.normalize_type == 1:
        return torch.sigmoid(raw)
    if args.normalize_type == 2:
        return torch.tanh(raw)
    assert False, 'Wrong normalize type'


def weights_init_normal(m):
    """Initialiser the reference applies when activation == 'relu' (models/helpers.py:63-71,139-140): N(0, 0.02) for
    convolution weights, N(1, 0.02) / 0 for BatchNorm2d.  It matches by class name, so the Linear layers of NPP-Net keep
    PyTorch's default initialisation -- applying it to the fused networks is a no-op, as it is in the reference."""
    kind = type(m).__name__
    if "Conv" in kind:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm2d" in kind:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0.0)


def create_npp_net(args, selected_angles, selected_periods, res, percep_net, is_search=False, style_net=None):
    """Same 7-tuple as the reference (models/helpers.py:75-175).  The model is never wrapped in nn.DataParallel:
    this framework runs one process per GPU (data parallelism is an NCCL all-reduce of the gradient arena)."""
    embedder, freq_nerf = get_embedder(args.multires, args.i_embed, res, is_search=is_search)
    if is_search:
        # one candidate periodicity, no top-K in the model (helpers.py:91-103)
        embedder_periodics, input_ch_periodic = get_embedder(
            args.multires, args.i_embed, res, selected_angles=selected_angles, selected_periods=selected_periods,
            freq_scales=args.freq_scales, freq_offsets=args.freq_offsets, angle_offsets=args.angle_offsets,
            is_search=True)
        model = NPP_Net_light(D=args.netdepth, W=args.netwidth, input_ch=freq_nerf, input_ch_periodic=input_ch_periodic,
                              freq_scales=args.freq_scales, freq_offsets=args.freq_offsets,
                              angle_offsets=args.angle_offsets, output_ch=3, skips=[4], activation=args.activation)
    else:
        embedder_periodics, input_ch_periodics = [], []
        for i in range(args.p_topk):
            e, ch = get_embedder(args.multires, args.i_embed, res, selected_angles=selected_angles[i],
                                 selected_periods=selected_periods[i], freq_scales=args.freq_scales,
                                 freq_offsets=args.freq_offsets, angle_offsets=args.angle_offsets)
            embedder_periodics.append(e)
            input_ch_periodics.append(ch)
        input_ch_periodics = np.array(input_ch_periodics)
        common = dict(D=args.netdepth, W=args.netwidth, freq_nerf=freq_nerf, freq_scales=args.freq_scales,
                      freq_offsets=args.freq_offsets, angle_offsets=args.angle_offsets, output_ch=3, skips=[4],
                      activation=args.activation)
        if args.p_topk > 1:
            model = NPP_Net(input_ch_periodic=input_ch_periodics[:1].sum(),
                            input_ch_periodic_aux=input_ch_periodics[1:].sum(), **common)
        else:
            model = NPP_Net_top1(input_ch_periodic=input_ch_periodics[:1].sum(), **common)

    if args.activation == 'relu':
        model.apply(weights_init_normal)
    grad_vars = list(model.parameters())
    if adaptive_pix is not None:
        grad_vars += list(adaptive_pix.parameters())
    if percep_net is not None and getattr(args, "use_adaptive_perceptual_loss", False):
        for adaptive in percep_net.adaptive_perceps:
            grad_vars += list(adaptive.parameters())
    if style_net is not None and getattr(args, "use_adaptive_style_loss", False):
        for adaptive in style_net.adaptives:
            grad_vars += list(adaptive.parameters())

    network_query_fn = lambda inputs, inputs_periodic, network_fn: run_network(  # noqa: E731
        inputs, inputs_periodic, network_fn, netchunk=args.netchunk)
    optimizer = NppAdam(grad_vars, lr=args.lrate, betas=(0.9, 0.999), net=model)
    start = 0
    render_kwargs_train = {'network_query_fn': network_query_fn, 'network_fn': model}
    render_kwargs_test = dict(render_kwargs_train)
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, embedder, embedder_periodics
