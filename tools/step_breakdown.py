"""Whole-step wall/device time vs the sum of profiled kernel classes (what is left is cuDNN parsing net,
allocator and launch gaps)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from archs.pgtformer_arch import PGTFormer  # noqa: E402
from pgtformer_b200 import ops  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kw = dict(bench.load_network_g())
kw.pop('type')
m = PGTFormer(**kw).cuda()
m.eval()
x = torch.rand(clips * 3, 3, size, size).cuda()
for _ in range(3):
    m(x, w=1, adain=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
t0 = time.perf_counter()
e0.record()
for _ in range(5):
    m(x, w=1, adain=True)
e1.record()
torch.cuda.synchronize()
print('step: %.2f ms device, %.2f ms wall  (%.1f clips/s)' % (e0.elapsed_time(e1) / 5, (time.perf_counter() - t0) * 200, clips * 5000 / e0.elapsed_time(e1)))
eng = m.engine()
e0.record()
for _ in range(5):
    eng.parse_pos(x)
e1.record()
torch.cuda.synchronize()
print('parse_pos (BiSeNet via cuDNN + convpos): %.2f ms' % (e0.elapsed_time(e1) / 5))
ops.profile_begin()
m(x, w=1, adain=True)
prof = ops.profile_end()
tot = 0
for k, v in prof.items():
    if v[2]:
        print('%-14s %8.2f ms %4d launches  %8.1f G(work)/ms' % (k, v[1], v[2], v[0] / v[1] / 1e9))
        tot += v[1]
print('sum profiled %.2f ms' % tot)
