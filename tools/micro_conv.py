"""Micro-benchmark of single conv / linear shapes (CUDA events, 20 reps)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgtformer_b200 import ops  # noqa: E402
from pgtformer_b200.engine import _pack_conv  # noqa: E402

dev = 'cuda'


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(F, H, C, N, residual, out_dtype=torch.bfloat16, act=0):
    x = torch.randn(F, H, H, C, device=dev).bfloat16()
    w = _pack_conv(torch.randn(N, C, 3, 3, device=dev) * 0.05)
    b = torch.zeros(N, device=dev)
    res = torch.randn(F, H, H, N, device=dev).to(out_dtype) if residual else None
    out = torch.empty(F, H, H, N, device=dev, dtype=out_dtype)
    ms = timeit(lambda: ops.conv(x, w, N, out, bias=b, residual=res, act=act))
    fl = 2.0 * F * H * H * N * 9 * C
    print('conv F%d H%d C%d N%d res=%d %s act=%d: %.3f ms  %.0f TF/s' % (F, H, C, N, residual, str(out_dtype)[6:], act, ms, fl / ms / 1e9))


for args in [(12, 512, 64, 64, False), (12, 512, 64, 64, True), (12, 512, 128, 64, False), (12, 256, 128, 128, False),
             (12, 256, 128, 128, True), (12, 128, 256, 256, True), (12, 128, 256, 256, False)]:
    conv_case(*args)
os.environ['PGT_NO_HALO'] = '1'
