import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgtformer_b200 import ops
T = 49152
g = torch.Generator().manual_seed(2)
cb = torch.randn(1025, 512, generator=g).cuda()
z = torch.randn(T, 512, generator=g).cuda().contiguous()
pack = ops.codebook_pack(cb, 1024)
idx = torch.empty(T, dtype=torch.int64, device='cuda')
ops.l2_argmin_tc(z, cb, pack, 1024, idx, None)
torch.cuda.synchronize()
ws = ops._last_argmin_ws
nfb = int(ws[0].item())
fb = ws[2:2 + nfb].tolist()
print('fallbacks', nfb)
zb = z.bfloat16().float(); cbb = cb[:1024].bfloat16().float()
dt = pack[1][:1024][None] - 2 * (zb @ cbb.t())
emax = pack[1][1024].sqrt(); demax = pack[1][1025].sqrt()
zn = z.norm(dim=1); dz = (z - zb).norm(dim=1)
D = 2 * (dz * emax + zn * demax) + zn * emax / 4096
W = 2 * D
lists = ws[2 + ((T + 1) // 2) * 2:].view(torch.float32).view(-1, 2, 16, 2)
for t in fb[:6]:
    srt = dt[t].sort().values
    inwin = int((dt[t] <= srt[0] + W[t]).sum())
    print('token', t, 'W %.3f' % W[t].item(), 'sigma %.3f' % dt[t].std().item(), 'in-window', inwin, 'best5', [round(v, 3) for v in srt[:5].tolist()])
    for gg in range(2):
        l = lists[t, gg]
        print('   list g%d' % gg, [(round(a, 2), int(torch.tensor(b).view(torch.int32))) for a, b in l.tolist()])
