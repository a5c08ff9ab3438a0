"""Env-gated run trace (the reference never prints its loss): APH_TRACE=<file.json> makes sim_func
record every similarity value (one D2H read per call -- only when tracing) and, at interpreter exit, writes
{sims, encode_image_calls, launches, wall_s}. Used by the script-level tests and the logs under profiles/."""
import atexit
import json
import os
import time

PATH = os.environ.get('APH_TRACE')
_state = {'sims': [], 'encodes': 0, 't0': time.time()}


def enabled():
    return PATH is not None


def sim(value):
    if PATH is not None:
        _state['sims'].append(float(value))


def encode():
    _state['encodes'] += 1


def _dump():
    from . import _lib
    out = {'sims': _state['sims'], 'encode_image_calls': _state['encodes'], 'wall_s': time.time() - _state['t0'],
           'launches': int(_lib.lib().aph_launch_count()) if _lib._lib is not None else 0}
    with open(PATH, 'w') as f:
        json.dump(out, f)


if PATH is not None:
    atexit.register(_dump)
