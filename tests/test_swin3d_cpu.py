"""Video-Swin BasicLayer: the oracle restatement against fixtures minted from the reference's own
`modules/swin.py` (oracle/make_golden.py --swin), and against the live reference module when /root/reference is present."""
import os

import pytest
import torch

from conftest import load_golden
from oracle import swin3d_oracle as S


def _oracle_layer(case):
    from modules.swin import BasicLayer                   # this repo's drop-in: used here only for its state-dict layout
    c = S.SWIN_CASES[case]
    shell = BasicLayer(c['dim'], c['depth'], c['heads'], c['window'])
    sd = S.synth_state(shell.state_dict(), c['seed'])
    return c, sd


@pytest.mark.parametrize('case', ['a', 'b'])
def test_oracle_matches_reference_golden(case):
    c, sd = _oracle_layer(case)
    g = load_golden('swin3d_%s.pt' % case)
    with torch.no_grad():
        y = S.basic_layer(sd, '', S.case_input(case), c['depth'], c['heads'], c['window'])
    assert y.shape == g['out'].shape
    assert (y - g['out'].float()).abs().max().item() < 1.5e-3 * g['out_absmax']      # fp16 storage of the fixture


def test_dropin_state_dict_is_reference_compatible():
    """Same parameter / buffer names and shapes as the reference's BasicLayer (checked against the live module when the
    reference tree is present, against the recorded count otherwise)."""
    from modules.swin import BasicLayer
    ours = BasicLayer(256, 4, 8, (5, 5, 5)).state_dict()
    assert len(ours) == 52 and ours['blocks.1.attn.relative_position_bias_table'].shape == (729, 8)
    from oracle.reference_loader import REFERENCE_ROOT, _ensure_paths, reference_available
    if not reference_available():
        pytest.skip('reference tree not present')
    import importlib.util
    _ensure_paths()
    spec = importlib.util.spec_from_file_location('_pgt_reference.modules.swin', os.path.join(REFERENCE_ROOT, 'modules', 'swin.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.BasicLayer(256, 4, 8, (5, 5, 5)).state_dict()
    assert set(ours) == set(ref)
    assert all(ours[k].shape == ref[k].shape and ours[k].dtype == ref[k].dtype for k in ref)
    assert torch.equal(ours['blocks.0.attn.relative_position_index'], ref['blocks.0.attn.relative_position_index'])
