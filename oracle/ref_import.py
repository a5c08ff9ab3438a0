"""Import the *real* reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE. Used by tests/golden/make_golden.py to generate the committed golden fixtures
and by CPU tests that are skipped when /root/reference is absent (it does not exist on the GPU box).

The reference hard-codes `.cuda()` and imports third-party modules that are absent in this image
(imageio, matplotlib, kornia, pywt, pytorch_wavelets, ipywidgets, IPython). We register empty stub
modules for those and make `.cuda()` the identity when no GPU is present, then load the reference's
`aphantasia` package under the private name `ref_aphantasia` so it can coexist with the drop-in.
Nothing is copied: the files are executed where they lie.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get('APH_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'aphantasia', 'image.py'))


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], '__aph_stub__', False):
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__aph_stub__ = True
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with .image, .utils, .transforms = the reference's own modules."""
    if 'ns' in _loaded:
        return _loaded['ns']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    _stub('imageio', imread=None, imsave=None)
    mp = _stub('matplotlib'); mp.pyplot = _stub('matplotlib.pyplot')
    for n in ('kornia', 'kornia.geometry', 'kornia.geometry.transform', 'kornia.filters'):
        _stub(n)
    _stub('kornia.filters.sobel', spatial_gradient=None)
    _stub('pywt'); _stub('pytorch_wavelets', DWTForward=None, DWTInverse=None)
    _stub('ipywidgets'); _stub('IPython')
    if not torch.cuda.is_available():  # the reference hard-codes .cuda()
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    pkg_dir = os.path.join(REF_ROOT, 'aphantasia')
    # The reference modules import each other as `aphantasia.xxx`; temporarily alias the package name.
    saved = {k: v for k, v in sys.modules.items() if k == 'aphantasia' or k.startswith('aphantasia.')}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType('aphantasia'); pkg.__path__ = [pkg_dir]
    sys.modules['aphantasia'] = pkg
    try:
        mods = {}
        for name in ('utils', 'transforms', 'image'):
            spec = importlib.util.spec_from_file_location('aphantasia.' + name, os.path.join(pkg_dir, name + '.py'))
            m = importlib.util.module_from_spec(spec)
            sys.modules['aphantasia.' + name] = m
            spec.loader.exec_module(m)
            setattr(pkg, name, m)
            mods[name] = m
    finally:
        for k in [k for k in sys.modules if k == 'aphantasia' or k.startswith('aphantasia.')]:
            sys.modules['ref_' + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    ns = types.SimpleNamespace(**mods)
    _loaded['ns'] = ns
    return ns
