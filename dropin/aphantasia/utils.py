"""Drop-in for /root/reference/aphantasia/utils.py -> aphantasia_b200.utils."""
from aphantasia_b200.utils import *  # noqa: F401,F403
import aphantasia_b200.utils as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
