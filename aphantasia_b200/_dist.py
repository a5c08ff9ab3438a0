"""One process per GPU (torchrun): lazy NCCL init, seed agreement, crop sharding, gradient all-reduce.

The reference is single-process and never seeds its RNGs. With N ranks every rank must replay the SAME
host random stream (so the crop table is identical and results do not depend on N) and start from the
same parameters: rank 0 draws one seed, broadcasts it, and every rank seeds torch + numpy with it.
"""
import os

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
        _state['rank'], _state['world'] = dist.get_rank(), dist.get_world_size()
        if os.environ.get('APH_SYNC_SEED', '1') == '1':
            dev = torch.device('cuda') if backend_is_nccl() else torch.device('cpu')
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64).to(dev)
            dist.broadcast(seed, 0)
            torch.manual_seed(int(seed.item())); np.random.seed(int(seed.item()) % (2 ** 32))
    _state['init'] = True
    return _state


def backend_is_nccl():
    import torch.distributed as dist
    return dist.is_initialized() and dist.get_backend() == 'nccl'


def rank():
    return init()['rank']


def world():
    return init()['world']


def all_reduce_sum_(t):
    """In-place SUM all-reduce over all ranks (NCCL over NVLink on GPU tensors)."""
    if world() > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
