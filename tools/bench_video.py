"""Streaming restoration throughput (frames/s, wall clock incl. H2D/D2H of rgb24 frames): the reference's one-window-per-
call loop shape (clips_per_batch=1), batched windows, and batched windows with per-frame work computed once."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from archs.pgtformer_arch import PGTFormer  # noqa: E402
from pgtformer_b200.video import VideoRestorer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kw = dict(bench.load_network_g())
kw.pop('type')
m = PGTFormer(**kw).cuda()
m.eval()
frames = np.random.RandomState(0).randint(0, 256, size=(n, size, size, 3), dtype=np.uint8)
for label, batch, reuse in (('one window per call', 1, False), ('16 windows per call', 16, False),
                            ('16 windows per call + per-frame reuse', 16, True)):
    vr = VideoRestorer(m, clips_per_batch=batch, reuse_frames=reuse)
    vr.restore(frames[:min(n, 2 * batch + 1)])                    # warm-up (allocator, lazy module state)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = vr.restore(frames)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%-42s %4d frames %dx%d: %7.1f frames/s  (%.2f ms/frame)' % (label, n, size, size, n / dt, 1e3 * dt / n))

# device time of one batch of 16 windows: clip-major input vs distinct frames + frame_index
eng = m.engine()
x18 = torch.rand(18, 3, size, size, device='cuda')
idx = torch.tensor([j for i in range(1, 17) for j in (i - 1, i, i + 1)], dtype=torch.int32, device='cuda')
x48 = x18[idx.long()].contiguous()
for label, fn in (('forward(48 frames)', lambda: eng.forward(x48, w=1.0, adain=True)),
                  ('forward(18 distinct frames, frame_index)', lambda: eng.forward(x18, w=1.0, adain=True, frame_index=idx))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    print('%-44s %.2f ms device, %.2f ms host enqueue' % (label, e0.elapsed_time(e1) / 5, 1e3 * host))
