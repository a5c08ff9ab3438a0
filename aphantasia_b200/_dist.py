"""One process per GPU (torchrun): lazy NCCL init, seed agreement, crop sharding, gradient all-reduce.

The reference is single-process and never seeds its RNGs. With N ranks every rank must replay the SAME
host random stream (so the crop table is identical and results do not depend on N) and start from the
same parameters: rank 0 draws one seed, broadcasts it, and every rank seeds torch + numpy with it.
"""
import atexit
import os
import sys

import numpy as np
import torch

_state = {'init': False, 'rank': 0, 'world': 1}


def init():
    if _state['init']:
        return _state
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group(backend=backend)
            atexit.register(_shutdown)          # a script that knows nothing about ranks (clip_fft.py) never destroys the group itself
        _state['rank'], _state['world'] = dist.get_rank(), dist.get_world_size()
        if os.environ.get('APH_SYNC_SEED', '1') == '1':
            dev = torch.device('cuda') if backend_is_nccl() else torch.device('cpu')
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64).to(dev)
            dist.broadcast(seed, 0)
            torch.manual_seed(int(seed.item())); np.random.seed(int(seed.item()) % (2 ** 32))
    _state['init'] = True
    return _state


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dist.destroy_process_group()
    except Exception:
        pass


def backend_is_nccl():
    import torch.distributed as dist
    return dist.is_initialized() and dist.get_backend() == 'nccl'


def rank():
    return init()['rank']


def world():
    return init()['world']


# ---- the path's one exchange: all-reduce of the canvas gradient. Default on NCCL ranks: our own NVLS / peer-memory kernel
# (csrc/comm.cu) on a buffer in symmetric memory; APH_COLLECTIVE=nccl (or any failure to set symmetric memory up) -> dist.all_reduce.
_sym = {'bufs': {}, 'mode': None, 'err': None}


def collective_mode():
    """'nvls' | 'p2p' | 'nccl' | 'gloo' | 'none' -- what all_reduce_sum_ uses for symmetric buffers."""
    return _sym['mode'] or ('none' if world() == 1 else ('nccl' if backend_is_nccl() else 'gloo'))


def symm_empty(shape):
    """A persistent fp32 CUDA tensor of `shape` in symmetric memory (two per shape, handed out alternately), or None when the
    symmetric path is unavailable. Collective: every rank must call it in the same order."""
    if world() == 1 or not backend_is_nccl() or os.environ.get('APH_COLLECTIVE', 'sym') == 'nccl' or _sym['mode'] == 'nccl':
        return None
    key = tuple(int(s) for s in shape)
    ent = _sym['bufs'].get(key)
    if ent is None:
        try:
            import ctypes as C
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            ring = []
            for _ in range(2):
                t = symm.empty(key, dtype=torch.float32, device=torch.device('cuda', torch.cuda.current_device()))
                h = symm.rendezvous(t, dist.group.WORLD)
                mc = int(h.multicast_ptr or 0) if os.environ.get('APH_COLLECTIVE', 'sym') != 'p2p' else 0
                peers = (C.c_uint64 * h.world_size)(*[int(p) for p in h.buffer_ptrs])
                ring.append({'t': t, 'h': h, 'mc': mc, 'peers': peers, 'pads': int(h.signal_pad_ptrs_dev)})
            if _sym['err'] is None:
                _sym['err'] = torch.zeros(1, dtype=torch.int32, device='cuda')
            _sym['mode'] = 'nvls' if ring[0]['mc'] else 'p2p'
            ent = _sym['bufs'][key] = {'ring': ring, 'next': 0}
            if rank() == 0:
                sys.stderr.write(' [aphantasia_b200] gradient exchange: own %s all-reduce kernel over symmetric memory (%d ranks)\n' % (_sym['mode'], world()))
        except Exception as ex:           # symmetric memory / multicast not available on this box: NCCL does the exchange
            _sym['mode'] = 'nccl'
            if rank() == 0:
                sys.stderr.write(' [aphantasia_b200] symmetric memory unavailable (%r): gradient exchange through NCCL all_reduce\n' % (ex,))
            return None
    slot = ent['ring'][ent['next']]
    ent['next'] ^= 1
    return slot['t']


def _sym_slot(t):
    for ent in _sym['bufs'].values():
        for slot in ent['ring']:
            if slot['t'].data_ptr() == t.data_ptr() and slot['t'].numel() == t.numel():
                return slot
    return None


def all_reduce_sum_(t):
    """In-place SUM all-reduce over all ranks: our NVLS / peer-memory kernel when `t` is a symmetric buffer (symm_empty), else
    torch.distributed (NCCL over NVLink on GPU tensors, gloo on CPU tensors)."""
    if world() > 1:
        slot = _sym_slot(t) if (t.is_cuda and _sym['bufs']) else None
        if slot is not None:
            from ._lib import check, lib, stream_ptr
            check(lib().aph_allreduce_sym(slot['mc'], slot['peers'], slot['pads'], rank(), world(), t.numel(), _sym['err'].data_ptr(), stream_ptr()),
                  'aph_allreduce_sym')
        else:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def collective_error():
    """True if a rank barrier of the symmetric all-reduce ever timed out (synchronises)."""
    return bool(_sym['err'] is not None and int(_sym['err'].item()) != 0)
