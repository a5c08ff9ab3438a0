"""Drop-in `clip` package -> aphantasia_b200.clip."""
from aphantasia_b200.clip import *  # noqa: F401,F403
from aphantasia_b200.clip import load, tokenize, available_models  # noqa: F401
