"""Deterministic inputs and weights for the golden cases.

Shared by ``make_golden.py`` (run once in the build container, where
``/root/reference`` exists) and by the tests (run anywhere).  Everything is
drawn from ``numpy.random.RandomState`` (legacy MT19937: bit-stable across
machines and numpy versions), so only the reference's OUTPUTS need to be stored
in the ``.npz`` fixtures.
"""
from __future__ import annotations

import numpy as np
import torch

E, H, D, ATT, KC, KS, VGG = 50, 64, 128, 64, 120, 3, 1000

# name -> dict(mode, B, S, L, S_ui, L_ui, V, Pc, vocab, seed, m_scale)
CASES = {
    "umpr_r":       dict(review_net_only=True,  B=5, S=4, L=6, S_ui=2, L_ui=5, V=1, Pc=1, vocab=43, seed=11, m_scale=1.0),
    "umpr_r_softM": dict(review_net_only=True,  B=3, S=3, L=5, S_ui=2, L_ui=4, V=1, Pc=1, vocab=31, seed=12, m_scale=0.05),
    "umpr_full_v1": dict(review_net_only=False, B=4, S=3, L=6, S_ui=3, L_ui=5, V=1, Pc=1, vocab=37, seed=13, m_scale=0.05),
    "umpr_full_v4": dict(review_net_only=False, B=6, S=4, L=7, S_ui=2, L_ui=6, V=4, Pc=2, vocab=41, seed=14, m_scale=1.0),
}
RNN_CASES = {
    # heavy ties; total_length (L) larger than the longest sequence; one-token sentences
    "rnn_ties":  dict(N=70, L=9, max_len=7, seed=21),
    "rnn_big":   dict(N=160, L=12, max_len=12, seed=22),
}


def gru_param_shapes(prefix):
    out = {}
    for suf in ("", "_reverse"):
        out[f"{prefix}.weight_ih_l0{suf}"] = (3 * H, E)
        out[f"{prefix}.weight_hh_l0{suf}"] = (3 * H, H)
        out[f"{prefix}.bias_ih_l0{suf}"] = (3 * H,)
        out[f"{prefix}.bias_hh_l0{suf}"] = (3 * H,)
    return out


def param_shapes(review_net_only: bool, V: int, vocab: int):
    """state_dict keys/shapes of the reference UMPR, in reference order."""
    s = {"embedding.weight": (vocab, E)}
    s.update(gru_param_shapes("review_net.r_net.gru.module"))
    s["review_net.r_net.M"] = (D, D)
    # nn.Module registers parameters (M) before sub-modules (gru) in named_parameters order; order is irrelevant here
    s["review_net.s_net_u.Ms"] = (ATT, D)
    s["review_net.s_net_u.Ws"] = (1, ATT)
    s["review_net.s_net_i.Ms"] = (ATT, D)
    s["review_net.s_net_i.Ws"] = (1, ATT)
    s["review_net.linear_u.weight"] = (D, 2 * D)
    s["review_net.linear_i.weight"] = (D, 2 * D)
    if not review_net_only:
        s.update(gru_param_shapes("control_net.c_net.gru.module"))
        s["control_net.c_net.cnn.0.weight"] = (KC, D, KS)
        s["control_net.c_net.cnn.0.bias"] = (KC,)
        s["control_net.c_net.linear.0.weight"] = (V, KC)
        s["control_net.c_net.linear.0.bias"] = (V,)
        s["control_net.s_net.Ms"] = (ATT, D)
        s["control_net.s_net.Ws"] = (1, ATT)
        s["control_net.ss_net.linear.0.weight"] = (1, D)
        s["control_net.ss_net.linear.0.bias"] = (1,)
        s["visual_net.pos_v_emb"] = (V, VGG)
        s["visual_net.neg_v_emb"] = (V, VGG)
        s["visual_net.linear.weight"] = (1, VGG)
        s["visual_net.linear.bias"] = (1,)
        s["linear_fusion.0.weight"] = (1, D + 2 * V)
    else:
        s["linear_fusion.0.weight"] = (1, D)
    s["linear_fusion.0.bias"] = (1,)
    return s


def make_params(review_net_only: bool, V: int, vocab: int, seed: int, m_scale: float = 1.0, dtype=torch.float32):
    """Weights with the reference's init *distributions* (not its RNG stream)."""
    rs = np.random.RandomState(seed)
    out = {}
    for k, shp in param_shapes(review_net_only, V, vocab).items():
        if k == "embedding.weight":
            a = rs.normal(0.0, 0.5, size=shp)
            a[:3] = 0.0                                    # word2vec.py:19-20 <PAD>/<UNK>/<NUM> are zero vectors
        elif ".gru.module." in k:
            a = rs.uniform(-1.0, 1.0, size=shp) / np.sqrt(H)
        elif k.endswith(".M"):
            a = rs.normal(size=shp) * m_scale
        elif k.endswith(".Ms") or k.endswith(".Ws"):
            a = rs.normal(size=shp) * 0.3
        elif k.endswith("v_emb"):
            a = rs.normal(size=shp)
        elif k == "linear_fusion.0.bias":
            a = np.full(shp, 3.0)                          # dead-ReLU guard (SURVEY.md §7 hazards)
        elif k == "control_net.c_net.linear.0.bias":
            a = np.linspace(-0.3, 0.3, shp[0]) if shp[0] > 1 else np.zeros(shp)
        else:
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            if k.endswith("bias"):
                fan_in = {"control_net.c_net.cnn.0.bias": D * KS, "control_net.ss_net.linear.0.bias": D,
                          "visual_net.linear.bias": VGG}.get(k, fan_in)
            a = rs.uniform(-1.0, 1.0, size=shp) / np.sqrt(fan_in)
            if k == "control_net.c_net.linear.0.weight":
                a = a * 6.0                                # spread sigmoid outputs around the 0.35 threshold
            if k == "control_net.ss_net.linear.0.weight":
                a = a * 40.0                               # push Eq.17 scores away from 0.5 so Eq.18's gates open
        out[k] = torch.tensor(np.asarray(a), dtype=dtype)
    return out


def make_lengths(rs, B, S, L, min_real=2):
    """(B,S) lengths >= 1 in the collate's style: some real sentences, the rest 1-token pads (dataset.py:125-127)."""
    lens = np.ones((B, S), dtype=np.int64)
    for b in range(B):
        c = rs.randint(1, S + 1)
        lens[b, :c] = rs.randint(min(min_real, L), L + 1, size=c)
    return lens


def make_reviews(rs, lens, L, vocab):
    B, S = lens.shape
    ids = np.zeros((B, S, L), dtype=np.int64)
    for b in range(B):
        for s in range(S):
            n = lens[b, s]
            if n > 1 or rs.rand() < 0.5:                  # a length-1 slot is either a real word or an all-PAD sentence
                ids[b, s, :n] = rs.randint(3, vocab, size=n)
    return ids


def make_batch(c):
    rs = np.random.RandomState(c["seed"] + 1000)
    B, S, L = c["B"], c["S"], c["L"]
    u_len = make_lengths(rs, B, S, L)
    i_len = make_lengths(rs, B, S, L)
    ui_len = make_lengths(rs, B, c["S_ui"], c["L_ui"])
    user = make_reviews(rs, u_len, L, c["vocab"])
    item = make_reviews(rs, i_len, L, c["vocab"])
    ui = make_reviews(rs, ui_len, c["L_ui"], c["vocab"])
    if c["review_net_only"]:
        photos = torch.zeros(0)                            # dataset.py:158,180 → Tensor([])
    else:
        photos = torch.tensor(rs.normal(0, 0.05, size=(B, c["V"], c["Pc"], VGG)), dtype=torch.float32)
    labels = torch.tensor(rs.randint(1, 6, size=B), dtype=torch.float32)
    t = lambda a: torch.tensor(a, dtype=torch.int64)
    return (t(user), t(item), t(ui), t(u_len), t(i_len), t(ui_len), photos, labels)


def make_rnn_case(c):
    rs = np.random.RandomState(c["seed"])
    N, L = c["N"], c["L"]
    lens = rs.randint(1, c["max_len"] + 1, size=N).astype(np.int64)
    data = rs.normal(0, 0.5, size=(N, L, E))
    for n in range(N):
        data[n, lens[n]:] = 0.0
    w = {k: torch.tensor(rs.uniform(-1, 1, size=s) / np.sqrt(H), dtype=torch.float32)
         for k, s in gru_param_shapes("module").items()}
    cot_out = torch.tensor(rs.normal(size=(N, L, D)), dtype=torch.float32)
    cot_hid = torch.tensor(rs.normal(size=(2, N, H)), dtype=torch.float32)
    return torch.tensor(data, dtype=torch.float32), torch.tensor(lens), w, cot_out, cot_hid


COLLATE_CASES = {"collate_a": dict(B=7, seed=31, max_count=6, max_ui=3, max_len=9), "collate_b": dict(B=33, seed=32, max_count=20, max_ui=5, max_len=20)}


def make_collate_case(c):
    """Ragged samples in the shape ``Dataset.__getitem__`` hands to ``batch_loader`` (dataset.py:38-41,153): (user sentences, item
    sentences, user->item sentences, photo ids (V, Pc), rating); sentences are token-id lists, some of them EMPTY (they become the
    1-token all-PAD sentences of dataset.py:125-127), pools are shorter than ``max_count`` for most samples."""
    rs = np.random.RandomState(c["seed"])

    def pool(max_count):
        n = rs.randint(1, max_count + 1)
        return [[int(t) for t in rs.randint(3, 500, size=rs.randint(0, c["max_len"] + 1))] for _ in range(n)]

    out = []
    for b in range(c["B"]):
        photos = [["p%d" % rs.randint(0, 40)] for _ in range(1)]
        out.append((pool(c["max_count"]), pool(c["max_count"]), pool(c["max_ui"]), photos, float(rs.randint(1, 6))))
    # the reference takes max_len over the raw sentence lengths: make sure at least one sentence per batch is non-empty
    out[0][0][0] = [int(t) for t in rs.randint(3, 500, size=c["max_len"])]
    return out


class CaseConfig:
    """The reference reads these attributes off ``config`` (model.py:235-253)."""
    def __init__(self, c):
        self.review_net_only = c["review_net_only"]
        self.loss_v_rate = 0.1
        self.gru_size = H
        self.self_atte_size = ATT
        self.views = ["v%d" % i for i in range(c["V"])]
        self.kernel_count = KC
        self.kernel_size = KS
        self.threshold = 0.35
