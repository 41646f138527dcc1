"""Runs a few representative ops in isolation (for `ncu --set full -k regex:...`)."""
import sys
import os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgtformer_b200 import ops  # noqa: E402
from pgtformer_b200.engine import _pack_conv  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'halo64'
dev = 'cuda'
F = 12
torch.manual_seed(0)
if which in ('halo64', 'halo128', 'conv256'):
    C, N, H = {'halo64': (64, 64, 512), 'halo128': (128, 128, 256), 'conv256': (256, 256, 128)}[which]
    x = torch.randn(F, H, H, C, device=dev).bfloat16()
    w = _pack_conv(torch.randn(N, C, 3, 3, device=dev) * 0.05)
    b = torch.zeros(N, device=dev)
    res = torch.randn(F, H, H, N, device=dev).bfloat16()
    out = torch.empty(F, H, H, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv(x, w, N, out, bias=b, residual=res)
elif which == 'rgb':
    from pgtformer_b200.engine import _pack_rgb
    x = torch.rand(F, 3, 512, 512, device=dev)
    w = _pack_rgb(torch.randn(64, 3, 3, 3, device=dev) * 0.1)
    b = torch.zeros(64, device=dev)
    out = torch.empty(F, 512, 512, 64, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(F * 2048 * 4 * 64, device=dev)
    for _ in range(3):
        ops.conv_rgb(x, w, b, out, 3, 1, 1, gn_stats=stats)
elif which == 'gn':
    x = torch.randn(F, 512, 512, 64, device=dev).bfloat16()
    g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.groupnorm_silu(x, g, b, out)
elif which == 'linear256':
    M = 786432 // 4
    a = torch.randn(M, 256, device=dev).bfloat16()
    w = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
    b = torch.zeros(256, device=dev)
    res = torch.randn(M, 256, device=dev).bfloat16()
    out = torch.empty(M, 256, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(a, w, out, bias=b, residual=res)
elif which == 'swin_mlp':
    M = 786432 // 4
    x = torch.randn(M, 256, device=dev).bfloat16()
    w1 = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
    w2 = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
    g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.swin_mlp(x, g, b, w1, b, w2, b, out)
torch.cuda.synchronize()
print('done', which)
