"""Drop-in for /root/reference/aphantasia/image.py -> aphantasia_b200.image."""
from aphantasia_b200.image import *  # noqa: F401,F403
import aphantasia_b200.image as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
