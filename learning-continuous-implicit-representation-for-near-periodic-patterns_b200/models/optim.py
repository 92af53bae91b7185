"""Optimizer returned by create_npp_net: torch.optim.Adam semantics (reference models/helpers.py:164) with the
NPP-Net parameters updated by one fused kernel over the plan arena and every foreign parameter (adaptive_pix,
LPIPS adaptive heads, ...) by a plain torch.optim.Adam.  ``param_groups[i]['lr']`` may be rewritten between steps
exactly as the reference loop does (NPP_completion/train.py:258-263)."""
import torch


class NppAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, net=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.net = net
        own = set(id(p) for p in net._params) if net is not None else set()
        self._own = [p for p in params if id(p) in own]
        foreign = [p for p in params if id(p) not in own]
        self._foreign = torch.optim.Adam(foreign, lr=lr, betas=betas, eps=eps) if foreign else None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        if self.net is not None and any(p.grad is not None for p in self._own):
            plan = self.net._plan
            for p, g in zip(self.net._params, self.net._grad_view_list):
                if g is None or p.grad is None or p.grad is g:
                    continue
                g.copy_(p.grad)                              # autograd cloned or accumulated: copy back into the arena
            plan.adam_step(group['lr'], betas=group['betas'], eps=group['eps'])
            self.net.mark_clean()
        if self._foreign is not None:
            for g in self._foreign.param_groups:
                g['lr'] = group['lr']
            self._foreign.step()
        return loss
