"""Drop-in for /root/reference/aphantasia/transforms.py -> aphantasia_b200.transforms."""
from aphantasia_b200.transforms import *  # noqa: F401,F403
import aphantasia_b200.transforms as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
