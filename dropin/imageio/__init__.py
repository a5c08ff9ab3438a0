"""Minimal imageio stand-in (absent in this image; clip_fft.py:6 imports imread/imsave at top level). PIL-backed."""
import numpy as np
from PIL import Image


def imread(path):
    return np.asarray(Image.open(path))


def imsave(path, img, **kw):
    Image.fromarray(np.asarray(img)).save(path, quality=95) if str(path).lower().endswith(('.jpg', '.jpeg')) else Image.fromarray(np.asarray(img)).save(path)


imwrite = imsave
