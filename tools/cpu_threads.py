import os, sys, time, torch
sys.path.insert(0, '.')
import bench
from oracle import pgt_oracle as O
from pgtformer_b200.spec import build_spec
from pgtformer_b200.weights import synth_state_dict
arch, spec = build_spec(bench.load_network_g())
sd = synth_state_dict(spec, 0)
x = torch.rand(3, 3, 256, 256)
print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for nt in (16, 32, 64, 128):
    torch.set_num_threads(nt)
    with torch.no_grad():
        O.pgtformer_forward(sd, arch, x, 1.0, True)
        t = time.perf_counter(); O.pgtformer_forward(sd, arch, x, 1.0, True); dt = time.perf_counter() - t
    print(nt, 'threads: %.2f s per 256^2 clip' % dt, flush=True)
