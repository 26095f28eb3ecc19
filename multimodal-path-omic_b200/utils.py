"""Helpers shared by the parameter containers (reference: models/utils.py)."""
import math

import torch.nn as nn


def init_max_weights(module):
    """N(0, 1/sqrt(fan_in)) weights and zero biases for every nn.Linear -- reference: models/utils.py:43-48
    (nn.Bilinear keeps torch's default init, exactly as there: the type check is `type(m) == nn.Linear`)."""
    for m in module.modules():
        if type(m) == nn.Linear:
            std = 1.0 / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, std)
            m.bias.data.zero_()
