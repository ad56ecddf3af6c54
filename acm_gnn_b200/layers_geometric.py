"""Drop-in replacement of ACM-Geometric's ``layers`` module (ACM-Geometric/layers.py:13-163).

Same class as the ACM-Pytorch flavour; the one behavioural difference is quirk Q1: here the
per-channel LayerNorm of the attention logits is LIVE for model types "acmgcnp"/"acmgcnpp"
(ACM-Geometric/layers.py:59,67), while ACM-Pytorch compares against "acmgcn+"/"acmgcn++"
which its CLI never passes.
"""
from .layers import MLP, device  # noqa: F401
from .layers import GraphConvolution as _Base


class GraphConvolution(_Base):
    _FLAVOUR = "geometric"
