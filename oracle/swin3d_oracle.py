"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Video-Swin `BasicLayer` of the reference
(`modules/swin.py:326-405`, used by `TDRQVAE` at `archs/tdrqvae_arch.py:834-835,850,854`), functional on a state dict.
Pinned against the reference module itself (imported with the mmcv / basicsr / timm shims of oracle/shims) by
oracle/make_golden.py --swin -> tests/golden/swin3d_*.pt and tests/test_swin3d_cpu.py.

Restated: get_window_size (:70-84), window_partition / window_reverse (:38-64), compute_mask (:309-323),
WindowAttention3D.forward (:136-166, incl. the `relative_position_index[:N, :N]` slice), SwinTransformerBlock3D
forward_part1 / forward_part2 / forward (:214-271: LN -> zero pad -> roll -> partition -> attention -> reverse -> roll
back -> crop -> + shortcut -> LN -> Mlp (ratio 4, exact GELU) -> + x), BasicLayer.forward (:380-405)."""
import torch
import torch.nn.functional as F


def window_size_for(x_size, window, shift):
    ws, ss = list(window), list(shift)
    for i in range(3):
        if x_size[i] <= window[i]:
            ws[i] = x_size[i]
            ss[i] = 0
    return tuple(ws), tuple(ss)


def partition(x, ws):
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws[0] * ws[1] * ws[2], C)


def reverse(win, ws, B, D, H, W):
    x = win.view(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, D, H, W, -1)


def shift_mask(Dp, Hp, Wp, ws, ss):
    img = torch.zeros(1, Dp, Hp, Wp, 1)
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    mw = partition(img, ws).squeeze(-1)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def relative_position_index(window):
    cd, ch, cw = (torch.arange(n) for n in window)
    coords = torch.stack(torch.meshgrid(cd, ch, cw, indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window[0] - 1
    rel[:, :, 1] += window[1] - 1
    rel[:, :, 2] += window[2] - 1
    rel[:, :, 0] *= (2 * window[1] - 1) * (2 * window[2] - 1)
    rel[:, :, 1] *= 2 * window[2] - 1
    return rel.sum(-1)


def window_attention(sd, p, xw, heads, mask):
    B_, N, C = xw.shape
    qkv = F.linear(xw, sd[p + '.qkv.weight'], sd.get(p + '.qkv.bias')).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = sd[p + '.relative_position_index'][:N, :N].reshape(-1)
    bias = sd[p + '.relative_position_bias_table'][idx].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    out = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[p + '.proj.weight'], sd[p + '.proj.bias'])


def block(sd, p, x, heads, window, shift, mask_matrix):
    B, D, H, W, C = x.shape
    ws, ss = window_size_for((D, H, W), window, shift)
    h = F.layer_norm(x, (C,), sd[p + '.norm1.weight'], sd[p + '.norm1.bias'], 1e-5)
    pd, pb, pr = (ws[0] - D % ws[0]) % ws[0], (ws[1] - H % ws[1]) % ws[1], (ws[2] - W % ws[2]) % ws[2]
    h = F.pad(h, (0, 0, 0, pr, 0, pb, 0, pd))
    _, Dp, Hp, Wp, _ = h.shape
    shifted = any(i > 0 for i in ss)
    if shifted:
        h = torch.roll(h, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    aw = window_attention(sd, p + '.attn', partition(h, ws), heads, mask_matrix if shifted else None)
    h = reverse(aw, ws, B, Dp, Hp, Wp)
    if shifted:
        h = torch.roll(h, shifts=ss, dims=(1, 2, 3))
    h = h[:, :D, :H, :W, :]
    x = x + h
    m = F.layer_norm(x, (C,), sd[p + '.norm2.weight'], sd[p + '.norm2.bias'], 1e-5)
    m = F.linear(F.gelu(F.linear(m, sd[p + '.mlp.fc1.weight'], sd[p + '.mlp.fc1.bias'])), sd[p + '.mlp.fc2.weight'], sd[p + '.mlp.fc2.bias'])
    return x + m


def basic_layer(sd, p, x, depth, heads, window):
    """x: [B, C, D, H, W] -> same.  `p`: state-dict prefix of the BasicLayer ('' or e.g. 'tdswin_pre')."""
    pre = (p + '.') if p else ''
    B, C, D, H, W = x.shape
    shift = tuple(i // 2 for i in window)
    ws, ss = window_size_for((D, H, W), window, shift)
    x = x.permute(0, 2, 3, 4, 1)
    Dp, Hp, Wp = -(-D // ws[0]) * ws[0], -(-H // ws[1]) * ws[1], -(-W // ws[2]) * ws[2]
    mask = shift_mask(Dp, Hp, Wp, ws, ss)
    for i in range(depth):
        x = block(sd, '%sblocks.%d' % (pre, i), x, heads, window, (0, 0, 0) if i % 2 == 0 else shift, mask)
    return x.permute(0, 4, 1, 2, 3)


SWIN_CASES = {'a': dict(dim=256, depth=4, heads=8, window=(5, 5, 5), shape=(1, 256, 3, 16, 16), seed=21),
              'b': dict(dim=256, depth=2, heads=8, window=(5, 5, 5), shape=(2, 256, 7, 10, 10), seed=22)}


def synth_state(state_dict, seed):
    """Deterministic stand-in checkpoint for a BasicLayer (no trained TDRQVAE weights are reachable): a pure function
    of (name, shape, seed), applied to any module with the reference's state-dict names."""
    out = {}
    for i, (k, v) in enumerate(sorted(state_dict.items())):
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if k.endswith('norm1.weight') or k.endswith('norm2.weight'):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('.bias'):
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith('relative_position_bias_table'):
            out[k] = 0.3 * torch.randn(v.shape, generator=g)
        else:
            out[k] = torch.randn(v.shape, generator=g) / (v.shape[-1] ** 0.5)
    return out


def case_input(case):
    c = SWIN_CASES[case]
    return torch.randn(*c['shape'], generator=torch.Generator().manual_seed(c['seed'] + 100))
