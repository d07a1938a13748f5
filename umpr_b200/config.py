"""The reference ``config.py`` switches that the model reads (config.py:7-39), with the same names and defaults.

The argparse front-end of the reference (config.py:41-52) is host orchestration and out of scope; any object with
these attributes (including the reference's own ``Config``) can be handed to ``UMPR(config, word_emb)``.
"""
from __future__ import annotations


class Config:
    multi_gpu = True
    batch_size = 64
    learning_rate = 1e-6
    l2_regularization = 1e-3
    lr_decay = 0.99
    review_net_only = False
    max_sent_count = 20
    min_sent_count = 5
    max_ui_sent_count = 5
    max_sent_length = 20
    views = ['unknown']          # 1 view for amazon; ['food', 'inside', 'outside', 'drink'] for yelp
    photo_count = 1
    gru_size = 64
    self_atte_size = 64
    kernel_count = 120
    kernel_size = 3
    threshold = 0.35
    loss_v_rate = 0.1

    def __init__(self, **overrides):
        for k, v in overrides.items():
            if not hasattr(type(self), k):
                raise AttributeError(f"unknown config switch {k!r}")
            setattr(self, k, v)
        if self.gru_size != 64 or self.self_atte_size != 64 or self.kernel_size != 3 or self.kernel_count > 128:
            raise NotImplementedError("umpr_b200 is built for gru_size=64, self_atte_size=64, kernel_size=3, kernel_count<=128")

    def __str__(self):
        keys = [k for k in dir(self) if not k.startswith('_') and not callable(getattr(self, k))]
        return ''.join(f'{k} = {getattr(self, k)}\n' for k in keys)
