"""Drop-in for /root/reference/aphantasia/progress_bar.py -> aphantasia_b200.progress_bar."""
from aphantasia_b200.progress_bar import *  # noqa: F401,F403
import aphantasia_b200.progress_bar as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
