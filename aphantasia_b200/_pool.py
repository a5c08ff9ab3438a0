"""Persistent device buffers for the per-step tensors of the drop-in (crops, gradients, canvases).

/root/reference/clip_fft.py:285 calls torch.cuda.empty_cache() EVERY step: whatever the step allocated through torch's caching
allocator is cudaFree'd (a device synchronisation each) and cudaMalloc'ed again a moment later -- measured 4-11 ms per step
through the unmodified script, more than the whole GPU step. The drop-in therefore takes its large per-step tensors from this
pool: base buffers that are never returned to the caching allocator, handed out as views and recycled when nothing references
their storage any more (torch's own storage use-count: views, autograd's saved tensors and user variables all count), so a
tensor the caller keeps across steps is never overwritten.
"""
import torch

_bases = {}            # (dtype, numel, device index) -> [1-D base tensors]
_MAX_PER_KEY = 6
stats = {'hits': 0, 'misses': 0, 'fallbacks': 0}


_USE_COUNT = getattr(torch._C, '_storage_Use_Count', None)      # private torch API; without it the pool degrades to plain torch.empty


def _in_use(base):
    # references to the StorageImpl: the base tensor itself + the temporary Python storage wrapper made for this query
    return _USE_COUNT(base.untyped_storage()._cdata) > 2


def empty(shape, dtype=torch.float32, device=None):
    """torch.empty(shape) on the current CUDA device, from the pool."""
    n = 1
    for s in shape:
        n *= int(s)
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if _USE_COUNT is None:
        stats['fallbacks'] += 1
        return torch.empty(tuple(shape), dtype=dtype, device='cuda')
    key = (dtype, n, dev)
    lst = _bases.setdefault(key, [])
    for base in lst:
        if not _in_use(base):
            stats['hits'] += 1
            return base.view(tuple(shape))
    if len(lst) >= _MAX_PER_KEY or n == 0:
        stats['fallbacks'] += 1
        return torch.empty(tuple(shape), dtype=dtype, device='cuda')
    stats['misses'] += 1
    base = torch.empty(n, dtype=dtype, device='cuda')
    lst.append(base)
    return base.view(tuple(shape))


def clear():
    _bases.clear()
