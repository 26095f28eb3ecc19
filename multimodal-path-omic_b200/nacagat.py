"""Drop-in NarrowContextualAttentionGateTransformer (reference: models/nacagat/nacagat.py)."""
from . import slidepath
from ._survival_model import SurvivalModelBase
from .blocks import PreGatingContextualAttention


class NarrowContextualAttentionGateTransformer(SurvivalModelBase):
    """NaCAGaT.  Same constructor / forward(wsi, omics) / outputs / state_dict keys as
    models/nacagat/nacagat.py:9-138; the co-attention map [6,N] is always returned (nacagat.py:93,136)."""

    variant = slidepath.VARIANT_NACAGAT

    def _make_coattention(self, width):
        return PreGatingContextualAttention(embed_dim=width, num_heads=1)

    def forward(self, wsi, omics):
        return self._run(wsi, omics, want_map=True)
