"""CPU test: the C-ABI library builds, loads, and exports every symbol include/pgt_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'pgt_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pgt_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from pgtformer_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_status_strings_and_arg_validation():
    from pgtformer_b200 import _lib
    lib = _lib.load()
    assert lib.pgt_version() >= 100
    assert lib.pgt_strerror(0) == b'ok'
    assert b'invalid' in lib.pgt_strerror(-1)
    # null pointers are rejected before any CUDA call is made
    assert lib.pgt_linear_bf16(None, 8, None, 8, 1, 1, 1, None, None) == -1
    assert lib.pgt_l2_argmin(None, 1, 512, None, 1024, None, None, None) == -1
    try:
        _lib.check(-3)
        raise AssertionError('check() must raise')
    except RuntimeError as e:
        assert 'not covered' in str(e)


def test_torch_library_shim_registers_the_ops():
    """`torch.ops.pgt.*` (pgtformer_b200/csrc_torch/pgt_torch_ops.cpp) builds, loads and registers its schemas — no
    compute without a GPU."""
    import torch
    from pgtformer_b200 import torch_ops
    ns = torch_ops.load()
    for name in ('window_attention', 'mha_fwd', 'argmax_gather', 'codebook_pack', 'l2_argmin', 'linear'):
        op = getattr(ns, name)
        assert 'pgt::' + name in str(op.default._schema)
    with __import__('pytest').raises(Exception):              # CPU tensors: no kernel registered for that backend
        torch.ops.pgt.codebook_pack(torch.zeros(8, 8), 8)
