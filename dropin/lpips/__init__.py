"""lpips stand-in: clip_fft.py:22 imports it at top level but only uses it with --sync (off by default)."""


class LPIPS:
    def __init__(self, *a, **k):
        raise NotImplementedError('lpips is not installed in this environment (only needed for --sync)')
