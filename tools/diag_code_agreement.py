#!/usr/bin/env python
"""Which stage costs code agreement?  Runs the CPU oracle (test infrastructure) at a small size to get the reference
intermediates, then feeds the GPU global transformer with every mix of (reference | GPU) lq_feat and parsing term and
reports the agreement of the resulting argmax codes with the reference's.  Diagnostic only — not a product path.
    python tools/diag_code_agreement.py [--size 256]
"""
import argparse
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    a = ap.parse_args()
    from archs.pgtformer_arch import PGTFormer
    from oracle import pgt_oracle as O
    from pgtformer_b200.spec import build_spec
    from pgtformer_b200.weights import synth_state_dict
    opt = yaml.safe_load(open(os.path.join(ROOT, 'options', 'release_test_stage_IIII_dont_need_align_version.yml')))['network_g']
    arch, spec = build_spec(opt)
    sd = synth_state_dict(spec, 0)
    kw = dict(opt)
    kw.pop('type')
    model = PGTFormer(**kw).cuda()
    model.eval()
    eng = model.engine()
    H = a.size
    x = torch.rand(3, 3, H, H, generator=torch.Generator().manual_seed(3))
    torch.set_num_threads(min(os.cpu_count(), 32))
    with torch.no_grad():
        (out, logits, lq_nhwc), inter = O.pgtformer_forward(sd, arch, x, 1.0, True, return_intermediates=True)
    hh = H // 16
    T = 3 * hh * hh
    ref_logits = logits.reshape(T, -1)
    ref_codes = ref_logits.argmax(-1)
    top2 = ref_logits.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    print('reference margins: median %.4f  p10 %.4f  p1 %.4f  min %.5f' % (margin.median(), margin.quantile(0.1), margin.quantile(0.01), margin.min()))
    # reference lq / pos in the engine's row order (clip, frame, y, x)
    ref_lq = lq_nhwc.reshape(T, -1)                                            # [3, h, w, E] -> rows (f, y, x)
    ref_pos = inter['pos'].reshape(T, -1)                                      # [t*th*tw, b=1, E], frame-major already
    xg = x.cuda()
    with torch.no_grad():
        gpu_pos = eng.parse_pos(xg).float()
        h, feats = eng.encoder(xg)
        gpu_lq32 = eng._lin(h.view(T, -1), 'quant_conv', arch.embed_dim, out_dtype=torch.float32)
    print('lq_feat : max rel err %.3e  mean abs err %.3e (max|ref| %.3f)' % ((gpu_lq32.cpu() - ref_lq).abs().max() / ref_lq.abs().max(), (gpu_lq32.cpu() - ref_lq).abs().mean(), ref_lq.abs().max()))
    print('pos     : max rel err %.3e  mean abs err %.3e (max|ref| %.3f)' % ((gpu_pos.cpu() - ref_pos).abs().max() / ref_pos.abs().max(), (gpu_pos.cpu() - ref_pos).abs().mean(), ref_pos.abs().max()))
    for name_lq, lq in (('ref', ref_lq.cuda()), ('gpu', gpu_lq32)):
        for name_pos, pos in (('ref', ref_pos.cuda()), ('gpu', gpu_pos)):
            with torch.no_grad():
                lo = eng.global_transformer(lq.bfloat16().contiguous(), pos.bfloat16().contiguous(), 1).float().cpu()
            err = (lo - ref_logits).abs()
            agree = (lo.argmax(-1) == ref_codes).float().mean().item()
            print('lq=%s pos=%s : code agreement %.4f   logit err max %.4f mean %.5f' % (name_lq, name_pos, agree, err.max(), err.mean()))
    # the reference's own bf16 sensitivity: the oracle transformer on bf16-rounded reference inputs (fp32 math)
    with torch.no_grad():
        q = O.linear(sd, 'feat_emb', ref_lq.bfloat16().float()).view(3 * hh * hh, 1, -1)
        pp = ref_pos.bfloat16().float().view(3 * hh * hh, 1, -1)
        for i in range(arch.n_layers):
            q = O.transformer_sa_layer(sd, 'ft_layers.%d' % i, q, pp, arch.n_head)
        lo = torch.nn.functional.linear(O.layer_norm(sd, 'idx_pred_layer.0', q), sd['idx_pred_layer.1.weight']).reshape(T, -1)
    print('oracle transformer on bf16-rounded ref inputs (fp32 math): agreement %.4f  logit err max %.4f mean %.5f' % (
        (lo.argmax(-1) == ref_codes).float().mean().item(), (lo - ref_logits).abs().max(), (lo - ref_logits).abs().mean()))


if __name__ == '__main__':
    main()
