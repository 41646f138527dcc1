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
elif which in ('linear512', 'linear512_f32'):
    # the 32^2-level Swin / global-transformer linears: M = 49152 tokens, N = K = 512 (bf16 or fp32 residual stream)
    M = 49152
    dt = torch.float32 if which.endswith('f32') else torch.bfloat16
    a = torch.randn(M, 512, device=dev).bfloat16()
    w = (torch.randn(512, 512, device=dev) * 0.05).bfloat16()
    b = torch.zeros(512, device=dev)
    res = torch.randn(M, 512, device=dev).to(dt)
    out = torch.empty(M, 512, device=dev, dtype=dt)
    for _ in range(3):
        ops.linear(a, w, out, bias=b, residual=res)
elif which == 'up128':
    from pgtformer_b200.engine import _pack_up2x
    x = torch.randn(F, 256, 256, 128, device=dev).bfloat16()
    w = _pack_up2x(torch.randn(128, 128, 3, 3, device=dev) * 0.05)
    out = torch.empty(F, 512, 512, 128, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv_up2x(x, w, 128, out, bias=torch.zeros(128, device=dev))
elif which == 'swin_mlp':
    M = 786432 // 4
    x = torch.randn(M, 256, device=dev).bfloat16()
    w1 = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
    w2 = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
    g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.swin_mlp(x, g, b, w1, b, w2, b, out)
elif which == 'window_tc':
    from pgtformer_b200.weights import relative_position_index
    C, H, clips, heads = 256, 128, 4, 8
    T = clips * 3 * H * H
    qkv = torch.randn(T, 3 * C, device=dev).bfloat16()
    bias = (0.02 * torch.randn(245, heads, device=dev))[relative_position_index().view(-1).to(dev)].view(48, 48, heads)
    tab = ops.window_tables(bias.permute(2, 0, 1).contiguous())
    out = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.window_attention_tc(qkv, clips, H, H, C, heads, 2, tab, out)
elif which == 'mha_tc':
    clips, L, E = 4, 3072, 512
    q, k, v = (torch.randn(clips * L, E, device=dev).bfloat16() for _ in range(3))
    out = torch.empty(clips * L, E, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.mha(q, k, v, clips, L, 8, 64, out)
elif which == 'argmax':
    T = 49152
    logits = torch.randn(T, 1024, device=dev)
    cb = torch.randn(1025, 512, device=dev)
    idx = torch.empty(T, dtype=torch.int64, device=dev)
    quant = torch.empty(T, 512, device=dev)
    for _ in range(3):
        ops.argmax_gather(logits, cb, idx, quant)
elif which == 'l2_argmin':
    T = 49152
    cb = torch.randn(1025, 512, device=dev)
    z = torch.randn(T, 512, device=dev)
    idx = torch.empty(T, dtype=torch.int64, device=dev)
    pack = ops.codebook_pack(cb, 1024)
    for _ in range(3):
        ops.l2_argmin_tc(z, cb, pack, 1024, idx, None)
elif which == 'ln_linear':
    M = 786432 // 4
    x = torch.randn(M, 256, device=dev).bfloat16()
    w = (torch.randn(768, 256, device=dev) * 0.05).bfloat16()
    g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    bias = torch.zeros(768, device=dev)
    out = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.ln_linear(x, g, b, w, bias, out)
elif which == 'conv_out':
    x = torch.randn(F, 512, 512, 64, device=dev).bfloat16()
    g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    ab = ops.groupnorm_ab(x, g, b, torch.empty(F * 128, device=dev))
    w = _pack_conv(torch.randn(3, 64, 3, 3, device=dev) * 0.05)
    out = torch.empty(F, 3, 512, 512, device=dev)
    for _ in range(3):
        ops.conv_out_gn(x, ab, w, 3, torch.zeros(3, device=dev), out)
torch.cuda.synchronize()
print('done', which)
