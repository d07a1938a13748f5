"""CPU oracle for the UMPR review-network hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``umpr_b200/`` imports it, and
the product path raises if the CUDA library is missing.

What it is
----------
A functional (no ``nn.Module``) restatement, in plain PyTorch-on-CPU tensor
algebra, of the arithmetic that ``/root/reference/src/model.py`` performs on
the path named by BASELINE.json ``north_star``.  Each function cites the
reference lines it follows.  Parameters travel in a flat ``dict`` whose keys
are the reference ``state_dict`` keys so the same dict can be
loaded into the reference modules, the oracle and the CUDA drop-in.

Third-party arithmetic
----------------------
The GRU cell, ``pack_padded_sequence``/``pad_packed_sequence``, ``torch.sort``,
softmax and conv1d are not reference code; they live in PyTorch (this image:
2.11.0+cu128, ATen CPU; the reference's readme states pytorch 1.7, no lock
file).  Two GRU restatements are provided:

* ``gru_explicit``  – an independent, per-time-step masked loop that follows the
  published GRU equations (torch.nn.GRU docs: r, z, n gates, rows ordered
  [r; z; n], h' = (1-z)*n + z*h).  Used for parity.
* ``gru_packed_lib`` – the same library calls the reference makes at
  ``model.py:18-20`` (pack → ``torch._VF.gru`` → pad).  Used for the CPU
  baseline timing because it is what the reference actually executes.

Pinning
-------
The reference ships no tests, golden vectors or fixtures for this path
("parity unpinned by the reference's tests").  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD
CONTAINER: ``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/src/model.py`` on CPU/fp32, runs seeded cases and commits
inputs, weights, forward returns and parameter gradients as ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor

EQ18_EPS = 1e-4  # model.py:188 (the readme and north_star say 1e-6; the code is the oracle)


# ----------------------------------------------------------------------------
# packing order  (model.py:18 -> torch.nn.utils.rnn.pack_padded_sequence)
# ----------------------------------------------------------------------------
def sort_plan(lengths: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """The permutation ``pack_padded_sequence(enforce_sorted=False)`` derives.

    Follows torch/nn/utils/rnn.py: ``lengths, sorted_indices = torch.sort(lengths,
    descending=True)`` on the CPU copy of the lengths (model.py:18 ``lengths.cpu()``),
    ``unsorted_indices = invert_permutation(sorted_indices)``.  The sort is NOT
    stable so the only bit-exact restatement is the same call.
    Returns (sorted_lengths, sorted_indices, unsorted_indices), all int64 CPU.
    """
    lengths = lengths.detach().to("cpu", torch.int64).reshape(-1)
    if lengths.numel() and (int(lengths.min()) < 1):
        # pack_padded_sequence raises for any length <= 0
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element "
                           "in 'lengths' that is <= 0")
    sorted_len, sorted_idx = torch.sort(lengths, descending=True)
    unsorted = torch.empty_like(sorted_idx)
    unsorted[sorted_idx] = torch.arange(sorted_idx.numel(), dtype=torch.int64)
    return sorted_len, sorted_idx, unsorted


def gru_weights(params: Dict[str, Tensor], prefix: str) -> Sequence[Tensor]:
    """Flat weight list in nn.GRU order: fwd (w_ih, w_hh, b_ih, b_hh) then reverse."""
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    return [params[f"{prefix}.{n}"] for n in names] + [params[f"{prefix}.{n}_reverse"] for n in names]


def _gru_cell(x_t: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step, torch.nn.GRU equations (gate rows [r; z; n])."""
    H = h.shape[-1]
    gi = x_t @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def gru_explicit(data: Tensor, lengths: Tensor, w: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """Bidirectional varlen GRU in ORIGINAL row order, zero beyond each length.

    Independent restatement of what pack→GRU→pad (model.py:18-20) yields before
    the second un-sort: ``Y[n, t] = [h_fwd(t) ‖ h_bwd(t)]`` for ``t < len[n]``,
    0 elsewhere; ``hidden[0, n]`` = forward state at ``t = len[n]-1``,
    ``hidden[1, n]`` = backward state at ``t = 0``.
    """
    N, L, _ = data.shape
    H = w[1].shape[1]
    lengths = lengths.to(data.device).reshape(-1)
    h = data.new_zeros(N, H)
    fwd = []
    for t in range(L):
        m = (lengths > t).unsqueeze(1)
        h_new = _gru_cell(data[:, t], h, *w[0:4])
        h = torch.where(m, h_new, h)
        fwd.append(torch.where(m, h_new, torch.zeros_like(h_new)))
    h_f = h
    h = data.new_zeros(N, H)
    bwd = [None] * L
    for t in range(L - 1, -1, -1):
        m = (lengths > t).unsqueeze(1)
        h_new = _gru_cell(data[:, t], h, *w[4:8])
        h = torch.where(m, h_new, h)
        bwd[t] = torch.where(m, h_new, torch.zeros_like(h_new))
    y = torch.cat([torch.stack(fwd, 1), torch.stack(bwd, 1)], dim=-1)
    return y, torch.stack([h_f, h], 0)


def gru_packed_lib(data: Tensor, lengths: Tensor, w: Sequence[Tensor], train: bool = True) -> Tuple[Tensor, Tensor]:
    """Same quantity as ``gru_explicit`` through the library calls of model.py:18-20."""
    pk = torch.nn.utils.rnn.pack_padded_sequence(data, lengths.cpu(), batch_first=True, enforce_sorted=False)
    H = w[1].shape[1]
    hx = data.new_zeros(2, pk.batch_sizes[0].item(), H)
    out, hidden = torch._VF.gru(pk.data, pk.batch_sizes, hx, list(w), True, 1, 0.0, train, True)
    out = torch.nn.utils.rnn.PackedSequence(out, pk.batch_sizes, pk.sorted_indices, pk.unsorted_indices)
    y, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=data.shape[1])
    return y, hidden.index_select(1, pk.unsorted_indices)


def improved_rnn(data: Tensor, lengths: Tensor, w: Sequence[Tensor], impl: str = "explicit") -> Tuple[Tensor, Tensor]:
    """``ImprovedRnn.forward`` (model.py:12-21) including the SECOND un-sort.

    ``result[n] = Y[unsorted_indices[n]]`` (model.py:21 indexes the already
    un-sorted padded tensor with ``package.unsorted_indices`` again, SURVEY.md
    §0.1); ``hidden`` is un-sorted once, i.e. in original order.
    ``total_length`` is ``data.shape[1]`` (model.py:17,20), never the batch max.
    """
    _, _, unsorted = sort_plan(lengths)
    if impl == "explicit":
        y, hidden = gru_explicit(data, lengths, w)
    else:
        y, hidden = gru_packed_lib(data, lengths, w)
    return y[unsorted.to(y.device)], hidden


# ----------------------------------------------------------------------------
# R-Net: co-attention  (model.py:36-56)
# ----------------------------------------------------------------------------
# Arg-max routing supplied from outside (tests only).  max() is continuous but its gradient is not: two implementations whose
# inputs differ in the last bits pick different winners among near-tied candidates and then legitimately disagree on every
# gradient upstream.  With ``routed({...})`` active the oracle takes the positions chosen by the implementation under test,
# records how far below the true maximum they are (``margin``: must be within that implementation's element error), and
# back-propagates through exactly those entries - parity is then defined again.
_ROUTING = None


class routed:
    """``with routed({"coattn": [(arg_u, arg_i)], "cnet": [cidx_ui, cidx_user, cidx_item]}) as r: ...`` ; afterwards ``r.margin``
    (how far below the oracle's own maximum the supplied winners are, relative to the value range) and ``r.stats`` (per kind:
    ``total`` maxima, ``differ`` = supplied position is not the oracle's own arg-max, ``strict`` = ... and its value is strictly
    below the oracle's maximum, i.e. a near-tie decided differently rather than an exact tie).

    ``mode``:
      "routed"         values and gradient paths of the supplied positions                                   (default)
      "routed_masked"  the same, but no gradient flows through the entries where the supplied position differs from the
                       oracle's own arg-max
      "own_masked"     the oracle's OWN maxima, no gradient through those same entries
    The two masked modes differ only in forward values at the masked entries (by at most ``margin``): if they agree, the supplied
    routing and the oracle's routing are identical everywhere else."""

    def __init__(self, picks, mode: str = "routed"):
        assert mode in ("routed", "routed_masked", "own_masked")
        self.picks = {k: list(v) for k, v in picks.items()}
        self.mode = mode
        self.margin = {"coattn": 0.0, "cnet": 0.0}
        self.stats = {k: {"total": 0, "differ": 0, "strict": 0} for k in ("coattn", "cnet")}

    def __enter__(self):
        global _ROUTING
        _ROUTING = self
        return self

    def __exit__(self, *exc):
        global _ROUTING
        _ROUTING = None

    def take(self, kind):
        return self.picks[kind].pop(0) if self.picks.get(kind) else None

    def resolve(self, kind, top, got, differ):
        """Book-keeping + the value to continue with.  top/got: the oracle's own maxima / the values at the supplied positions."""
        st = self.stats[kind]
        st["total"] += differ.numel()
        st["differ"] += int(differ.sum())
        st["strict"] += int((differ & (got.detach() < top.detach())).sum())
        self.margin[kind] = max(self.margin[kind], float(((top - got).abs().max() / top.abs().max().clamp_min(1e-30)).detach()))
        if self.mode == "routed":
            return got
        val = got if self.mode == "routed_masked" else top
        return torch.where(differ, val.detach(), val)


def _routed_max(t: Tensor, dim: int, idx, kind: str) -> Tensor:
    top, own = t.max(dim=dim)
    if idx is None:
        return top
    idx = idx.to(torch.int64).clamp(min=0)
    got = t.gather(dim, idx.unsqueeze(dim)).squeeze(dim)
    return _ROUTING.resolve(kind, top, got, idx != own)


def co_attention(gru_u: Tensor, gru_i: Tensor, M: Tensor):
    """model.py:50-55.  Unmasked max / softmax over all P = S*L positions."""
    A = torch.tanh(gru_i @ M @ gru_u.transpose(-1, -2))
    pick = _ROUTING.take("coattn") if _ROUTING is not None else None
    soft_u = torch.softmax(_routed_max(A, -2, None if pick is None else pick[0], "coattn"), dim=-1)
    soft_i = torch.softmax(_routed_max(A, -1, None if pick is None else pick[1], "coattn"), dim=-1)
    atte_u = (gru_u.transpose(-1, -2) @ soft_u.unsqueeze(-1)).squeeze(-1)
    atte_i = (gru_i.transpose(-1, -2) @ soft_i.unsqueeze(-1)).squeeze(-1)
    return soft_u, soft_i, atte_u, atte_i


def r_net(user_emb: Tensor, item_emb: Tensor, u_len: Tensor, i_len: Tensor,
          params: Dict[str, Tensor], prefix: str = "review_net.r_net", impl: str = "explicit"):
    """``RNet.forward`` (model.py:36-56) → the 6-tuple that is API (pretrain_rnet.py:165)."""
    B, S, L, E = user_emb.shape
    w = gru_weights(params, f"{prefix}.gru.module")
    gru_u, _ = improved_rnn(user_emb.reshape(B * S, L, E), u_len.reshape(-1), w, impl)
    gru_i, _ = improved_rnn(item_emb.reshape(B * S, L, E), i_len.reshape(-1), w, impl)
    gru_u = gru_u.reshape(B, S * L, -1)
    gru_i = gru_i.reshape(B, S * L, -1)
    soft_u, soft_i, atte_u, atte_i = co_attention(gru_u, gru_i, params[f"{prefix}.M"])
    return gru_u, gru_i, soft_u, soft_i, atte_u, atte_i


# ----------------------------------------------------------------------------
# S-Net: sentence self-attention  (model.py:71-81)
# ----------------------------------------------------------------------------
def s_net(gru_repr: Tensor, word_soft: Tensor, sent_length: int, Ms: Tensor, Ws: Tensor):
    B = gru_repr.shape[0]
    S = gru_repr.shape[1] // sent_length
    x = gru_repr.reshape(B * S, sent_length, -1)                       # (N, L, D)
    score = torch.tanh(x @ Ms.t()) @ Ws.t()                            # (N, L, 1)   model.py:76
    sent_soft = torch.softmax(score.squeeze(-1), dim=-1)               # over L, unmasked
    self_atte = (sent_soft.unsqueeze(-1) * x).sum(1)                   # (N, D)      model.py:77
    w = word_soft.reshape(B * S, -1).sum(-1, keepdim=True)             # model.py:79
    sentiment = (w * self_atte).reshape(B, S, -1).sum(1)               # model.py:80
    return self_atte.reshape(B, S, -1), sentiment


# ----------------------------------------------------------------------------
# ReviewNet  (model.py:157-169)
# ----------------------------------------------------------------------------
def review_net(user_emb, item_emb, u_len, i_len, params, impl="explicit"):
    L = user_emb.shape[-2]
    gru_u, gru_i, soft_u, soft_i, atte_u, atte_i = r_net(user_emb, item_emb, u_len, i_len, params, impl=impl)
    _, senti_u = s_net(gru_u, soft_u, L, params["review_net.s_net_u.Ms"], params["review_net.s_net_u.Ws"])
    _, senti_i = s_net(gru_i, soft_i, item_emb.shape[-2], params["review_net.s_net_i.Ms"], params["review_net.s_net_i.Ws"])
    repr_u = torch.cat([atte_u, senti_u], -1)
    repr_i = torch.cat([atte_i, senti_i], -1)
    return torch.tanh(repr_u @ params["review_net.linear_u.weight"].t() + repr_i @ params["review_net.linear_i.weight"].t())


# ----------------------------------------------------------------------------
# C-Net  (model.py:110-126)
# ----------------------------------------------------------------------------
def c_net(review_emb, lengths, params, threshold, prefix="control_net.c_net", impl="explicit"):
    B, S, L, E = review_emb.shape
    w = gru_weights(params, f"{prefix}.gru.module")
    g, _ = improved_rnn(review_emb.reshape(B * S, L, E), lengths.reshape(-1), w, impl)
    gru_repr = g.reshape(B, S * L, -1)
    cw, cb = params[f"{prefix}.cnn.0.weight"], params[f"{prefix}.cnn.0.bias"]
    pad = (cw.shape[-1] - 1) // 2
    pre = torch.nn.functional.conv1d(g.transpose(-1, -2), cw, cb, padding=pad)                # (N, K, L) model.py:118
    pick = _ROUTING.take("cnet") if _ROUTING is not None else None
    if pick is None:
        feat = torch.relu(pre).max(dim=-1)[0]                                                 # model.py:119-121
    else:
        # routed: position AND the ReLU decision (index -1 = clipped) come from the implementation under test - relu(max) has a
        # kink at 0 just like max has one at a tie
        idx = pick.to(torch.int64)
        top_pre, own = pre.max(dim=-1)
        top = torch.relu(top_pre)
        own = torch.where(top_pre > 0, own, torch.full_like(own, -1))                         # the oracle's own routing, same encoding
        got = torch.where(idx >= 0, pre.gather(-1, idx.clamp(min=0).unsqueeze(-1)).squeeze(-1), torch.zeros_like(top))
        feat = _ROUTING.resolve("cnet", top, got, idx != own)
    feat = feat.reshape(B, S, -1)
    view_p = torch.sigmoid(feat @ params[f"{prefix}.linear.0.weight"].t() + params[f"{prefix}.linear.0.bias"])
    view_p = torch.where(view_p < threshold, torch.zeros_like(view_p), view_p)                # model.py:124
    final = (view_p ** 2).sum(-2)                                                             # model.py:125
    return gru_repr, view_p, final


# ----------------------------------------------------------------------------
# ControlNet  (model.py:179-198)  — Eq. 18 and the quadratic gates
# ----------------------------------------------------------------------------
def control_net(user_emb, item_emb, ui_emb, u_len, i_len, ui_len, params, threshold, eps=EQ18_EPS, impl="explicit"):
    L_ui = ui_emb.shape[-2]
    gru_repr, view_p, c_out = c_net(ui_emb, ui_len, params, threshold, impl=impl)
    _, _, c_u = c_net(user_emb, u_len, params, threshold, impl=impl)
    _, _, c_i = c_net(item_emb, i_len, params, threshold, impl=impl)
    s, _ = s_net(gru_repr, view_p, L_ui, params["control_net.s_net.Ms"], params["control_net.s_net.Ws"])
    senti = torch.sigmoid(s @ params["control_net.ss_net.linear.0.weight"].t() + params["control_net.ss_net.linear.0.bias"])
    p2 = view_p ** 2
    score = (senti * p2).sum(-2) / (p2.sum(-2) + eps)                                          # model.py:188
    q_p = (score > 0.5).to(score.dtype)                                                        # model.py:189,192
    q_pos = torch.where(score < 0.5, torch.zeros_like(score), 4 * (score - 0.5) ** 2)          # model.py:190,193
    q_neg = torch.where(score > 0.5, torch.zeros_like(score), 4 * (0.5 - score) ** 2)          # model.py:191,194
    prefer_pos = c_out * q_p * q_pos
    prefer_neg = c_out * (1 - q_p) * q_neg
    return c_u, c_i, prefer_pos, prefer_neg


# ----------------------------------------------------------------------------
# VisualNet tail  (model.py:218-229; the VGG16 backbone :204-207,217 is out of scope)
# ----------------------------------------------------------------------------
def visual_net_tail(features, c_u, c_i, params, prefix="visual_net"):
    """``features``: (B, V, Pc, 1000) — what ``self.vgg16(images).view(...)`` yields at model.py:218."""
    img = features.mean(dim=-2)
    lw, lb = params[f"{prefix}.linear.weight"], params[f"{prefix}.linear.bias"]
    img_emb = (img @ lw.t() + lb).squeeze(-1)
    pos_emb = (params[f"{prefix}.pos_v_emb"] @ lw.t() + lb).squeeze(-1)
    neg_emb = (params[f"{prefix}.neg_v_emb"] @ lw.t() + lb).squeeze(-1)
    pos_match = torch.tanh(torch.abs(pos_emb - img_emb))
    neg_match = torch.tanh(torch.abs(neg_emb - img_emb))
    return pos_match, neg_match, c_u * c_i * (1 - pos_match), c_u * c_i * (1 - neg_match)


# ----------------------------------------------------------------------------
# UMPR  (model.py:257-278)
# ----------------------------------------------------------------------------
def umpr_forward(params: Dict[str, Tensor], batch, *, review_net_only: bool, threshold: float = 0.35,
                 loss_v_rate: float = 0.1, eps: float = EQ18_EPS, impl: str = "explicit"):
    """Returns (prediction (B,), loss scalar).  ``batch`` is the 8-tuple of dataset.py:173-182;
    ``photos`` holds VGG16 *features* (B, V, Pc, 1000[,1,1])."""
    user_reviews, item_reviews, ui_reviews, u_len, i_len, ui_len, photos, labels = batch
    table = params["embedding.weight"]
    user_emb = table[user_reviews]                                                             # model.py:262-264
    item_emb = table[item_reviews]
    represent = review_net(user_emb, item_emb, u_len, i_len, params, impl=impl)
    fw, fb = params["linear_fusion.0.weight"], params["linear_fusion.0.bias"]
    if review_net_only:
        pred = torch.relu(represent @ fw.t() + fb).squeeze(-1)
        return pred, torch.mean((pred - labels) ** 2)
    ui_emb = table[ui_reviews]
    c_u, c_i, prefer_pos, prefer_neg = control_net(user_emb, item_emb, ui_emb, u_len, i_len, ui_len, params,
                                                   threshold, eps, impl=impl)
    feats = photos.reshape(photos.shape[0], photos.shape[1], photos.shape[2], -1)
    pos_match, neg_match, final_pos, final_neg = visual_net_tail(feats, c_u, c_i, params)
    pred = torch.relu(torch.cat([represent, final_pos, final_neg], -1) @ fw.t() + fb).squeeze(-1)
    loss_r = torch.mean((pred - labels) ** 2)
    loss_v = torch.mean(prefer_pos.t() @ pos_match + prefer_neg.t() @ neg_match)               # model.py:276
    return pred, loss_r + loss_v * loss_v_rate


# ----------------------------------------------------------------------------
# R-Net pre-training  (pretrain/pretrain_rnet.py:144-169) – next-row (f3)
# ----------------------------------------------------------------------------
def pretrain_rnet_forward(params: Dict[str, Tensor], u: Tensor, u_len: Tensor, i: Tensor, i_len: Tensor, target: Tensor,
                          impl: str = "explicit"):
    """``PretrainRNet.forward``: ids (B,L) -> one sentence per sample -> R-Net -> sigmoid(Linear(256->1)) -> BCELoss.
    Keys: ``embedding.weight``, ``r_net.*``, ``linear.0.{weight,bias}``."""
    table = params["embedding.weight"]
    ue = table[u.reshape(u.shape[0], 1, u.shape[1])]                                           # pretrain_rnet.py:158-163
    ie = table[i.reshape(i.shape[0], 1, i.shape[1])]
    out = r_net(ue, ie, u_len.reshape(-1, 1), i_len.reshape(-1, 1), params, prefix="r_net", impl=impl)
    att = torch.cat([out[4], out[5]], dim=-1)                                                  # :165
    result = torch.sigmoid(att @ params["linear.0.weight"].t() + params["linear.0.bias"]).squeeze(-1)   # :166
    return result, torch.nn.functional.binary_cross_entropy(result, target)                   # :167


# ----------------------------------------------------------------------------
# Collate  (src/dataset.py:122-131,153-182) and the photo features  (dataset.py:134-151, model.py:216-218) – next-rows (f2), (f4)
# ----------------------------------------------------------------------------
def pad_reviews(reviews, max_count=None, max_len=None, pad=0):
    """dataset.py:122-131: pad every sample to ``max_count`` sentences (missing ones are EMPTY lists) and every sentence to ``max_len``
    tokens with ``pad``; ``lengths = max(1, len)`` - an empty slot is a 1-token all-PAD sentence."""
    if max_count is None:
        max_count = max(len(r) for r in reviews)
    lengths = [[max(1, len(r[j])) if j < len(r) else 1 for j in range(max_count)] for r in reviews]
    if max_len is None:
        max_len = max(max(row) for row in lengths)
    ids = torch.full((len(reviews), max_count, max_len), pad, dtype=torch.int64)
    for b, r in enumerate(reviews):
        for j, sent in enumerate(r):
            ids[b, j, :len(sent)] = torch.tensor(sent, dtype=torch.int64)
    return ids, torch.tensor(lengths, dtype=torch.int64)


def batch_loader(batch_list, pad=0):
    """dataset.py:153-182 without the JPEG loading: user and item padded to SHARED (max_count, max_len) taken over the raw sentence
    lengths of both (:163-170), the user->item review to its own maxima (:171).  → (user, item, ui, u_len, i_len, ui_len, labels)."""
    ru, ri, rui = ([s[k] for s in batch_list] for k in range(3))
    max_count = max(max(len(a), len(b)) for a, b in zip(ru, ri))
    max_len = max(max(max(len(s) for s in a), max(len(s) for s in b)) for a, b in zip(ru, ri))
    u, ul = pad_reviews(ru, max_count, max_len, pad)
    i, il = pad_reviews(ri, max_count, max_len, pad)
    ui, uil = pad_reviews(rui, pad=pad)
    return u, i, ui, ul, il, uil, torch.tensor([s[4] for s in batch_list], dtype=torch.float32)


def photo_features(table: Tensor, rows: Tensor, missing_row: int = -1) -> Tensor:
    """What ``VisualNet`` sees at model.py:218 when the backbone's outputs are cached per photo: ``table[rows]``; a row < 0 (photo
    missing / unreadable: the reference feeds a zero image, dataset.py:147-148) takes ``table[missing_row]``, or zeros."""
    rows = rows.to(torch.int64)
    out = table[rows.clamp(min=0)]
    miss = table[missing_row] if missing_row >= 0 else torch.zeros(table.shape[1], dtype=table.dtype)
    return torch.where((rows < 0).unsqueeze(-1), miss.expand_as(out), out)


def trainable_keys(params: Dict[str, Tensor]):
    """Everything but the frozen embedding (model.py:237 ``from_pretrained`` ⇒ requires_grad=False)."""
    return [k for k in params if k != "embedding.weight"]


def umpr_loss_and_grads(params, batch, **kw):
    """forward + backward of the train-step body (main.py:33-36); returns pred, loss, {key: grad}."""
    p = {k: v.detach().clone().requires_grad_(k != "embedding.weight") for k, v in params.items()}
    pred, loss = umpr_forward(p, batch, **kw)
    keys = trainable_keys(p)
    grads = torch.autograd.grad(loss, [p[k] for k in keys], allow_unused=True)
    return pred.detach(), loss.detach(), {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(keys, grads)}


# ----------------------------------------------------------------------------
# Optimiser restatement (main.py:22-26,37,54) – next-row (f1)
# ----------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor update with L2 (weight_decay added to the gradient)."""
    g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
    return p
