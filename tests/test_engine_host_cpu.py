"""Host-side algebra of the engine that needs no GPU: the LayerNorm affine folded into the linear layer that follows it
(`pgtformer_b200/engine.py::fold_layernorm_affine`; norm1 -> q/kv and norm2 -> fc1 of the C = 256 Swin blocks,
`modules/rstt_layers.py:298-336`), checked against the unfolded composition in float64."""
import torch
import torch.nn.functional as F


def test_fold_layernorm_affine_is_exact_algebra():
    from pgtformer_b200.engine import fold_layernorm_affine
    g = torch.Generator().manual_seed(11)
    C, N, T = 256, 768, 37
    x = torch.randn(T, C, generator=g, dtype=torch.float64) * 3 + 0.7
    gamma = 1 + 0.2 * torch.randn(C, generator=g, dtype=torch.float64)
    beta = 0.3 * torch.randn(C, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, generator=g, dtype=torch.float64) / C ** 0.5
    c = torch.randn(N, generator=g, dtype=torch.float64)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + c
    wf, cf = fold_layernorm_affine(w.float(), c.float(), gamma.float(), beta.float())
    assert wf.dtype == torch.float32 and cf.dtype == torch.float32 and wf.shape == (N, C) and cf.shape == (N,)
    xh = F.layer_norm(x, (C,), None, None, 1e-5)
    got = xh @ wf.double().t() + cf.double()
    assert (got - ref).abs().max().item() < 2e-5 * ref.abs().max().item()      # fp32 rounding of the folded parameters only


def test_fold_layernorm_identity_affine_is_a_no_op():
    from pgtformer_b200.engine import fold_layernorm_affine
    g = torch.Generator().manual_seed(12)
    w, c = torch.randn(64, 256, generator=g), torch.randn(64, generator=g)
    wf, cf = fold_layernorm_affine(w, c, torch.ones(256), torch.zeros(256))
    assert torch.equal(wf, w) and torch.equal(cf, c)
