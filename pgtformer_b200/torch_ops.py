"""`torch.ops.pgt.*`: the TORCH_LIBRARY binding of the C ABI (pgtformer_b200/csrc_torch/pgt_torch_ops.cpp), next to the
ctypes one (pgtformer_b200/_lib.py).  Both call the same extern "C" entry points of libpgt_b200.so; the engine uses the
ctypes binding, this module is the PyTorch-native route (dispatcher-registered ops, usable from TorchScript / C++):

    from pgtformer_b200 import torch_ops; torch_ops.load()
    torch.ops.pgt.window_attention(qkv, clips, H, W, C, heads, shift, tab16, out)
    torch.ops.pgt.l2_argmin(z, codebook, cb16, norm, 1024, idx, None)
"""
import torch

_loaded = False


def load():
    """Builds (if needed) and loads lib/libpgt_torch.so; returns torch.ops.pgt."""
    global _loaded
    if not _loaded:
        from . import _lib, build
        _lib.load()                                   # libpgt_b200.so first (the shim links against it)
        torch.ops.load_library(build.build_torch_shim())
        _loaded = True
    return torch.ops.pgt
