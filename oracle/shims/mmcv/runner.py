"""Import shim (test infrastructure): `modules/swin.py:8` imports mmcv.runner.load_checkpoint, which only its
`init_weights` (never called on this path) uses."""


def load_checkpoint(*args, **kwargs):
    raise RuntimeError('mmcv is not installed: load_checkpoint is a stub of the oracle shims')
