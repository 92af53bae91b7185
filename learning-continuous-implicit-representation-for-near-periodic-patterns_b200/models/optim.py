"""Optimizer returned by create_npp_net: torch.optim.Adam semantics (reference models/helpers.py:164) with the
NPP-Net parameters updated by one fused kernel over the plan arena and every foreign parameter (adaptive_pix,
LPIPS adaptive heads, ...) by one small CUDA Adam launch each (npp_adam_flat; torch.optim.Adam for anything that is not a
dense fp32 CUDA tensor).  ``param_groups[i]['lr']`` may be rewritten between steps
exactly as the reference loop does (NPP_completion/train.py:258-263)."""
import torch

from ._core import native as _nat


class NppAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, net=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.net = net
        own = set(id(p) for p in net._params) if net is not None else set()
        self._own = [p for p in params if id(p) in own]
        foreign = [p for p in params if id(p) not in own]
        # dense fp32 CUDA tensors (every foreign parameter of the reference scripts): one launch each, same arithmetic
        self._small = [p for p in foreign if p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()]
        self._small_state = {}
        rest = [p for p in foreign if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous())]
        self._foreign = torch.optim.Adam(rest, lr=lr, betas=betas, eps=eps) if rest else None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if len(self.param_groups) != 1:
            raise RuntimeError("NppAdam updates the whole parameter arena with one set of hyper-parameters: "
                               "exactly one param_group is supported (the reference builds one, helpers.py:164)")
        group = self.param_groups[0]
        if self.net is not None and any(p.grad is not None for p in self._own):
            plan = self.net._plan
            # the fused kernel runs over the whole trained arena: a frozen tensor or one without a gradient would be
            # updated from stale arena contents where torch.optim.Adam skips it
            bad = [n for n, p, g in zip(self.net._param_names, self.net._params, self.net._grad_view_list)
                   if g is not None and (p.grad is None or not p.requires_grad)]
            if bad:
                raise RuntimeError("NppAdam cannot skip individual NPP-Net tensors (no gradient / requires_grad=False): "
                                   + ", ".join(bad[:4]) + (" ..." if len(bad) > 4 else ""))
            for p, g in zip(self.net._params, self.net._grad_view_list):
                if g is None or p.grad is None or p.grad is g:
                    continue
                g.copy_(p.grad)                              # autograd cloned or accumulated: copy back into the arena
            plan.adam_step(group['lr'], betas=group['betas'], eps=group['eps'])
            self.net.mark_clean()
        for p in self._small:
            if p.grad is None:
                continue                                     # torch.optim skips parameters without a gradient
            st = self._small_state.get(id(p))
            if st is None:
                st = self._small_state[id(p)] = [0, torch.zeros_like(p), torch.zeros_like(p)]
            st[0] += 1
            g = p.grad if p.grad.is_contiguous() and p.grad.dtype == torch.float32 else p.grad.contiguous().float()
            _nat.check(_nat.lib().npp_adam_flat(p.data_ptr(), g.data_ptr(), st[1].data_ptr(), st[2].data_ptr(), p.numel(),
                                                group['lr'], group['betas'][0], group['betas'][1], group['eps'], st[0],
                                                _nat.current_stream()))
        if self._foreign is not None:
            for g in self._foreign.param_groups:
                g['lr'] = group['lr']
            self._foreign.step()
        return loss
