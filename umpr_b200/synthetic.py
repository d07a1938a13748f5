"""Seeded synthetic inputs shaped like the reference's collate output (dataset.py:122-182).

Workloads follow SURVEY.md §8(d): GloVe-shaped table (rows 0-2 zero, word2vec.py:12-20), lengths >= 1
(dataset.py:127), user/item padded to a shared (S, L), the user→item review to its own (S_ui, L_ui), photos given as
VGG16 *features*.  Everything is generated on the host (CPU tensors), as ``batch_loader`` would.
"""
from __future__ import annotations

import numpy as np
import torch

WORKLOADS = {
    # name: review_net_only, S, L, S_ui, L_ui, views, photo_count, long_pools
    "music_small_r": dict(review_net_only=True, S=20, L=20, S_ui=5, L_ui=20, V=1, Pc=1, long_pools=False),   # configs[0]
    "music_full":    dict(review_net_only=False, S=20, L=20, S_ui=5, L_ui=20, V=1, Pc=1, long_pools=False),  # configs[1]
    "yelp_full":     dict(review_net_only=False, S=20, L=20, S_ui=5, L_ui=20, V=4, Pc=1, long_pools=False),  # configs[2]
    "csj_long":      dict(review_net_only=False, S=20, L=20, S_ui=5, L_ui=20, V=1, Pc=1, long_pools=True),   # configs[3]
}


def make_table(vocab: int = 400003, dim: int = 50, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(vocab, dim, generator=g) * 0.5
    t[:3] = 0.0
    return t


def _side(rs, B, S, L, vocab, min_count, long_pools, skew):
    lens = np.ones((B, S), dtype=np.int64)
    ids = np.zeros((B, S, L), dtype=np.int64)
    lo = min(6, L)
    for b in range(B):
        c = S if long_pools else rs.randint(min(min_count, S), S + 1)
        if skew:
            ll = np.where(rs.rand(c) < 0.1, L, rs.randint(1, min(4, L) + 1, size=c))
        elif long_pools:
            ll = np.where(rs.rand(c) < 0.7, L, rs.randint(lo, L + 1, size=c))
        else:
            ll = rs.randint(lo, L + 1, size=c)
        ll = -np.sort(-ll)                      # pools are sorted longest-first (dataset.py:69-71)
        lens[b, :c] = ll
    mask = np.arange(L)[None, None, :] < lens[:, :, None]
    real = np.zeros((B, S), dtype=bool)
    real[lens > 1] = True
    ids_all = rs.randint(3, vocab, size=(B, S, L))
    ids = np.where(mask & real[:, :, None], ids_all, 0)     # length-1 padding sentences are all-PAD (dataset.py:125-127)
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(lens)


def make_batch(workload: str, batch: int, vocab: int = 400003, seed: int = 0, *, L=None, skew=False):
    """→ the 8-tuple (user_reviews, item_reviews, ui_reviews, u_lengths, i_lengths, ui_lengths, photos, labels)."""
    w = dict(WORKLOADS[workload])
    if L is not None:
        w["L"] = L
        w["L_ui"] = L
    rs = np.random.RandomState(seed)
    user, ul = _side(rs, batch, w["S"], w["L"], vocab, 5, w["long_pools"], skew)
    item, il = _side(rs, batch, w["S"], w["L"], vocab, 5, w["long_pools"], skew)
    ui, uil = _side(rs, batch, w["S_ui"], w["L_ui"], vocab, 1, False, skew)
    if w["review_net_only"]:
        photos = torch.zeros(0)
    else:
        photos = torch.from_numpy(rs.normal(0, 0.05, size=(batch, w["V"], w["Pc"], 1000)).astype(np.float32))
    labels = torch.from_numpy(rs.randint(1, 6, size=batch).astype(np.float32))
    return user, item, ui, ul, il, uil, photos, labels


def workload_config(workload: str):
    from .config import Config
    w = WORKLOADS[workload]
    return Config(review_net_only=w["review_net_only"], views=["v%d" % i for i in range(w["V"])], photo_count=w["Pc"])


def build_model(workload: str, table: torch.Tensor, seed: int = 0, device="cuda"):
    """Reference-style init under a seed, plus the dead-ReLU guard of SURVEY.md §7 (fusion bias 3.0)."""
    from .model import UMPR
    torch.manual_seed(seed)
    m = UMPR(workload_config(workload), table)
    with torch.no_grad():
        m.linear_fusion[0].bias.fill_(3.0)
    return m.to(device)
