"""Shared body of the two survival models (MCAT and NaCAGaT differ only in their co-attention block)."""
import torch
import torch.nn as nn

from .blocks import AttentionNetGated
from .fusion import BilinearFusion, ConcatFusion, GatedConcatFusion
from . import slidepath

_WIDTH = {"small": 128, "medium": 256, "big": 512}


def _snn_block(d_in, d_out, p):
    return nn.Sequential(nn.Linear(d_in, d_out), nn.ELU(), nn.AlphaDropout(p=p, inplace=False))


class SurvivalModelBase(nn.Module):
    """Parameter layout of models/mcat/mcat.py:13-82 (identical in models/nacagat/nacagat.py:10-78), created in the
    reference's order with torch's own initialisers so that seeds and checkpoints line up."""

    variant = None          # set by subclasses
    always_returns_map = False

    def _make_coattention(self, width):
        raise NotImplementedError

    def __init__(self, omic_sizes, model_size='medium', n_classes=4, dropout=0.25, fusion='concat', device='cpu'):
        super().__init__()
        self.n_classes = n_classes
        if model_size in _WIDTH:
            self.model_sizes = [_WIDTH[model_size], _WIDTH[model_size]]
        w0, w1 = self.model_sizes      # AttributeError for an unknown size, as in the reference
        self.dropout = dropout
        self.omic_sizes = list(omic_sizes)

        self.H = nn.Sequential(nn.Linear(1024, w0), nn.ReLU(), nn.Dropout(dropout))
        self.G = nn.ModuleList([nn.Sequential(_snn_block(d, w0, dropout), _snn_block(w0, w1, dropout))
                                for d in omic_sizes])
        self.co_attention = self._make_coattention(w1)

        def encoder():
            layer = nn.TransformerEncoderLayer(d_model=w1, nhead=8, dim_feedforward=512, dropout=dropout,
                                               activation='relu')
            return nn.TransformerEncoder(layer, num_layers=2)

        def rho():
            return nn.Sequential(nn.Linear(w1, w1), nn.ReLU(), nn.Dropout(dropout))

        self.path_transformer = encoder()
        self.path_attention_head = AttentionNetGated(n_classes=1, input_dim=w1, hidden_dim=w1)
        self.path_rho = rho()
        self.omic_transformer = encoder()
        self.omic_attention_head = AttentionNetGated(n_classes=1, input_dim=w1, hidden_dim=w1)
        self.omic_rho = rho()

        self.fusion = fusion
        if fusion == 'concat':
            self.fusion_layer = ConcatFusion(dims=[w1, w1], hidden_size=w1, output_size=w1).to(device=device)
        elif fusion == 'bilinear':
            self.fusion_layer = BilinearFusion(dim1=w1, dim2=w1, output_size=w1)
        elif fusion == 'gated_concat':
            self.fusion_layer = GatedConcatFusion(dims=[w1, w1], hidden_size=w1, output_size=w1).to(device=device)
        else:
            raise RuntimeError(f'Fusion mechanism {self.fusion} not implemented')
        self.classifier = nn.Linear(w1, n_classes)
        self._engine_obj = None
        self.model_size = model_size
        # nn.DataParallel (models/mcat/main.py:267-268 wraps the model whenever the box has more than one GPU) calls
        # forward() on replicas whose _parameters dicts are EMPTY (the weights are plain broadcast tensors there).  A
        # replica shares this __dict__ entry, so it finds the module that owns the leaf parameters: the engine binds
        # to that one, and gradients land in its .grad exactly where DataParallel's reduction would put them.
        self._master_ref = (self,)

    @property
    def _engine(self):
        master = self._master_ref[0]
        if master is not self:
            return master._engine
        if self._engine_obj is None:
            if self.model_sizes[0] != 256:
                raise NotImplementedError(
                    "the B200 kernels are built for model_size='medium' (width 256); '%s' has no kernel and there "
                    "is no PyTorch fallback" % self.model_size)
            binding = slidepath.ModelBinding(self, self.variant, self.fusion, self.omic_sizes, self.n_classes)
            self._engine_obj = slidepath.SlideEngine(binding, bag_dropout=self.dropout)
        return self._engine_obj

    def _run(self, wsi, omics, want_map):
        if len(omics) != len(self.omic_sizes):
            raise RuntimeError("expected %d omic groups, got %d" % (len(self.omic_sizes), len(omics)))
        hazards, S, Y, coattn, a_path, a_omic = slidepath.run_slide(self._engine, wsi, list(omics), want_map,
                                                                    self.training)
        return hazards, S, Y, {'coattn': coattn, 'path': a_path, 'omic': a_omic}

    def get_trainable_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
