"""TEST INFRASTRUCTURE ONLY — mints tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/reference_loader.py) on seeded inputs with the
deterministic synthetic checkpoint (pgtformer_b200.weights.synth_state_dict, seed 0).

Run in the build container (the reference does not exist on the GPU box):
    python -m oracle.make_golden
Inputs are not stored: `golden_input(seed, b, H)` regenerates them bit-exactly.
"""
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
OPT = os.path.join(ROOT, 'options', 'release_test_stage_IIII_dont_need_align_version.yml')


def load_network_g():
    with open(OPT) as f:
        return yaml.safe_load(f)['network_g']


def golden_input(seed, b, H):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b * 3, 3, H, H, generator=g)


def main():
    from oracle.reference_loader import build_reference_model, reference_forward, import_reference, generalise_size
    from pgtformer_b200.spec import build_spec
    from pgtformer_b200.weights import synth_state_dict
    opt = load_network_g()
    arch, spec = build_spec(opt)
    sd = synth_state_dict(spec, 0)
    m = build_reference_model(opt, sd)
    ref_mod = import_reference()
    torch.set_num_threads(os.cpu_count())
    for (seed, b, H) in ((1, 1, 128), (2, 2, 128)):
        x = golden_input(seed, b, H)
        out, logits, lq = reference_forward(m, x, w=1.0, adain=True)
        rec = {'out': out, 'logits': logits, 'lq_feat': lq.contiguous(), 'seed': seed, 'b': b, 'H': H,
               'w': 1.0, 'adain': True}
        # the registered TDCRQVAE3.forward (L2-argmin path) on the same module / weights
        vq = []
        with torch.no_grad():
            for i in range(b):
                generalise_size(m, H, H)
                vq.append(ref_mod.TDCRQVAE3.forward(m, x[i * 3:(i + 1) * 3]))
        rec['vq_out'] = torch.cat([v[0] for v in vq], 0)
        rec['vq_loss'] = torch.stack([v[1] for v in vq])
        rec['vq_codes'] = torch.cat([v[2] for v in vq], 0)
        path = os.path.join(GOLDEN, 'pgtformer_ref_b%d_%d_seed%d.pt' % (b, H, seed))
        torch.save(rec, path)
        print('wrote', path, {k: tuple(v.shape) for k, v in rec.items() if torch.is_tensor(v)})


if __name__ == '__main__':
    main()
