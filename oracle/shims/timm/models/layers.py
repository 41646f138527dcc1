import torch.nn as nn


class DropPath(nn.Module):
    """Stochastic depth; identity in eval mode / p == 0 (the only regime the oracle uses)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not (self.training and self.drop_prob > 0.0), 'oracle shim: eval / p=0 only'
        return x


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def to_2tuple(x):
    return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)
