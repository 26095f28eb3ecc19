"""Drop-in MultimodalCoAttentionTransformer (reference: models/mcat/mcat.py) running on the B200 slide engine."""
import torch.nn as nn

from . import slidepath
from ._survival_model import SurvivalModelBase


class MultimodalCoAttentionTransformer(SurvivalModelBase):
    """MCAT.  Same constructor, forward signature, outputs and state_dict keys as models/mcat/mcat.py:12-142.

    forward(wsi, omics, inference=False) -> hazards [1,4], survs [1,4], Y [1,4],
    {'coattn': [6,N] when inference else None, 'path': [1,6], 'omic': [1,6]}.
    wsi: [N,1024] or [1,N,1024] (fp32 or bf16) on the GPU; omics: 6 tensors [d_i] or [1,d_i]."""

    variant = slidepath.VARIANT_MCAT

    def _make_coattention(self, width):
        # nn.MultiheadAttention is used purely as the parameter container (same names / init as mcat.py:48)
        return nn.MultiheadAttention(embed_dim=width, num_heads=1)

    def forward(self, wsi, omics, inference: bool = False):
        return self._run(wsi, omics, want_map=bool(inference))
