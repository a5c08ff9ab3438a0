"""Adam behind torch.optim's interface, running in libaphb200.so.

The reference builds `torch.optim.Adam(params, lr, betas=(.0, .999))` over the one spectrum tensor and calls
zero_grad() / loss.backward() / step() per iteration (/root/reference/clip_fft.py:108-115,293-295). `Adam` here keeps that
interface (param_groups with a mutable 'lr' -- the script's --prog schedule writes it, clip_fft.py:288-291 --, state_dict,
zero_grad, step) and does the update in ONE launch (aph_adam_step) instead of torch's ~12 foreach launches; when its parameter
is the spectrum of an `fft_image`, the update is fused into the last pass of the synthesis backward (aph_synth_fft_bwd_adam:
dP is in registers there), so `step()` launches nothing and dP is never written.

Opt-in for an unmodified script: `APH_FUSED_ADAM=1 python -m aphantasia_b200.run clip_fft.py ...` makes `torch.optim.Adam`
resolve to this class when (and only when) every parameter belongs to one of our generators and no unsupported option
(weight_decay, amsgrad, maximize) is set; anything else gets the stock optimiser.
"""
import weakref

import torch

from ._lib import check, lib, stream_ptr

_GENERATORS = weakref.WeakValueDictionary()        # id(param tensor) -> FFTImage


def register_generator(param, gen):
    _GENERATORS[id(param)] = gen


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unsupported):
        if weight_decay != 0 or amsgrad or any(unsupported.get(k) for k in ('maximize', 'capturable', 'differentiable')):
            raise NotImplementedError('aphantasia_b200.optim.Adam: plain Adam only (no weight_decay / amsgrad / maximize)')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.fused_steps = 0
        for group in self.param_groups:
            for p in group['params']:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError('aphantasia_b200.optim.Adam: parameters must be contiguous fp32 CUDA tensors')
                self.state[p] = {'step': 0, 'exp_avg': torch.zeros_like(p), 'exp_avg_sq': torch.zeros_like(p), 'fused_done': False}
                gen = _GENERATORS.get(id(p))
                if gen is not None and gen.params is p:
                    gen.fused_opt = (weakref.ref(self), group, p)

    # called from _SynthFFT.backward (image.py) when exactly one grad-tracked synthesis happened since the last step()
    def _fused_args(self, group, p):
        st = self.state[p]
        st['step'] += 1
        st['fused_done'] = True
        self.fused_steps += 1
        b1, b2 = group['betas']
        return st['exp_avg'], st['exp_avg_sq'], float(group['lr']), float(b1), float(b2), float(group['eps']), int(st['step'])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                st = self.state[p]
                gen = _GENERATORS.get(id(p))
                if gen is not None:
                    gen.pending_fwd = 0
                if st['fused_done']:                 # the update already happened inside the synthesis backward
                    st['fused_done'] = False
                    continue
                if p.grad is None:
                    continue
                st['step'] += 1
                g = p.grad.contiguous()
                check(lib().aph_adam_step(p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), p.numel(),
                                          float(group['lr']), float(b1), float(b2), float(group['eps']), int(st['step']), stream_ptr()), 'aph_adam_step')
        return loss


_stock_adam = torch.optim.Adam


def _adam_factory(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
    plist = list(params)
    plain = weight_decay == 0 and not amsgrad and not any(kw.get(k) for k in ('maximize', 'capturable', 'differentiable', 'foreach', 'fused'))
    ours = plain and len(plist) > 0 and all(isinstance(p, torch.Tensor) and id(p) in _GENERATORS for p in plist)
    if ours:
        return Adam(plist, lr, betas, eps)
    return _stock_adam(plist, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)


def install():
    """Routes torch.optim.Adam(...) over our generators' parameters to the fused implementation (APH_FUSED_ADAM=1)."""
    torch.optim.Adam = _adam_factory


def uninstall():
    torch.optim.Adam = _stock_adam
