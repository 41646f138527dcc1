import sys, os, torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgtformer_b200 import ops
from pgtformer_b200.spec import build_spec
from pgtformer_b200.weights import synth_state_dict
opt = yaml.safe_load(open(os.path.join(ROOT, 'options/release_test_stage_IIII_dont_need_align_version.yml')))['network_g']
arch, spec = build_spec(opt)
sd = synth_state_dict(spec, 0)
cbf = sd['quantizer.codebooks.0.weight'].clone()
def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed); return torch.randn(shape, generator=g) * scale
for T in (775, 128, 1, 5000):
    z = 30.0 * rnd((T, 512), 90)
    cbd = cbf.cuda().contiguous()
    pack = ops.codebook_pack(cbd, 1024)
    idx = torch.full((T,), -7, dtype=torch.int64, device='cuda')
    ops.l2_argmin_tc(z.cuda().contiguous(), cbd, pack, 1024, idx, None)
    torch.cuda.synchronize()
    ws = ops._last_argmin_ws
    nfb = int(ws[0].item())
    fb = set(ws[2:2 + nfb].tolist())
    d64 = ((z.double().cuda()[:, None, :] - cbd[:1024].double()[None]) ** 2).sum(-1)
    ref = d64.argmin(1)
    bad = (idx != ref).nonzero().flatten().tolist()
    print('T', T, 'fallbacks', nfb, 'mismatches', len(bad))
    zb = z.cuda().bfloat16().float(); cbb = cbd[:1024].bfloat16().float()
    dt = pack[1][:1024][None] - 2 * (zb @ cbb.t())
    for t in bad[:5]:
        g, r = idx[t].item(), ref[t].item()
        order = dt[t].argsort()
        print('  token', t, 'in fallback list', t in fb, 'got', g, 'ref', r, 'd64 got %.6f ref %.6f' % (d64[t, g].item(), d64[t, r].item()),
              'approx rank of got/ref', (order == g).nonzero().item(), (order == r).nonzero().item(),
              'approx d got %.3f ref %.3f min %.3f' % (dt[t, g].item(), dt[t, r].item(), dt[t].min().item()))
        lst = ws[2 + ((T + 1) // 2) * 2:].view(torch.float32).view(-1, 2, 16, 2)[t]
        print('   lists g0', [(round(a, 2), int(torch.tensor(b).view(torch.int32))) for a, b in lst[0].tolist()][:8])
        print('   lists g1', [(round(a, 2), int(torch.tensor(b).view(torch.int32))) for a, b in lst[1].tolist()][:8])
