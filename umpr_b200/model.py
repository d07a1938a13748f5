"""Drop-in mirror of the reference ``src/model.py`` module tree on the B200 kernels.

Same class names, constructor arguments, ``forward`` signatures, return tuples and ``state_dict`` keys
(SURVEY.md §8b), so ``main.py:33`` / ``evaluate.py:11`` / ``pretrain_rnet.py:165`` style host code works unchanged and reference
checkpoints load through ``load_reference_checkpoint`` (it drops the out-of-scope ``visual_net.vgg16.*`` backbone keys).  The ``nn`` sub-modules (``nn.GRU``, ``nn.Conv1d``, ``nn.Linear``) are kept ONLY as
parameter containers with the reference's initialisation; their ``forward`` is never called — every operator runs
through ``umpr_b200.functional`` → ``libumpr_b200.so``.  There is no CPU path: parameters must be on a CUDA device.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as F
from .plan import PackPlan

EQ18_EPS = 1e-4   # model.py:188 (the code, not the readme's 1e-6, is the oracle)
NATIVE_EVAL = True  # UMPR.forward under no_grad goes through the one-call native step (csrc/step.cu) when the batch fits its envelope


def _on(device):
    """Every kernel is launched on the CURRENT device's stream (``_lib.call``): make the device that holds the data current for the
    duration of a forward (the reference lets the model live on any ``config.device``).  Backward needs no guard: the autograd engine
    runs each node on the worker thread of the device its forward ran on."""
    if device.type != "cuda":
        raise RuntimeError("umpr_b200: move the module to a CUDA device first (there is no CPU path)")
    return torch.cuda.device(device)


def _load_pretrained(module, path, what):
    """``pretrained=`` of model.py:31-34,66-69,101-104,135-138: the reference pickles whole modules (``torch.save(model)``) and swallows
    any failure after printing.  Same behaviour, except that the load uses ``weights_only=False`` (PyTorch >= 2.6 would otherwise
    reject every such pickle and the pre-trained weights would silently never load) and that a failure is also a RuntimeWarning."""
    import warnings
    try:
        obj = torch.load(path, weights_only=False)
        module.load_state_dict(obj.state_dict() if hasattr(obj, "state_dict") else obj)
    except Exception as e:                                   # noqa: BLE001 - the reference's bare except
        print(f'Failed to load {what} pre-trained weights from "{path}"')
        warnings.warn(f'umpr_b200: {what} pre-trained weights NOT loaded from "{path}": {e!r}', RuntimeWarning, stacklevel=3)


def load_reference_checkpoint(model, checkpoint, strict: bool = True):
    """Load a checkpoint written by the reference (``torch.save(model)`` of a ``src.model.UMPR``, main.py:47-52 - or its state_dict)
    into the drop-in ``model``.  The reference's ``visual_net.vgg16.*`` backbone parameters have no counterpart here (the backbone is
    upstream of this path, ``photos`` are its features) and are dropped; every other key must match when ``strict``."""
    obj = torch.load(checkpoint, weights_only=False) if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "read") else checkpoint
    sd = obj.state_dict() if hasattr(obj, "state_dict") else dict(obj)
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}          # a DataParallel wrapper (main.py:82)
    sd = {k: v for k, v in sd.items() if not k.startswith("visual_net.vgg16.")}
    return model.load_state_dict(sd, strict=strict)


class PackedReviews:
    """One review side (user, item or user→item) ready for the GRU kernels: plan + packed inputs, built once and
    shared by every ImprovedRnn that reads the same tokens (R-Net's and C-Net's GRUs, model.py:45-46,182-184)."""

    def __init__(self, lengths, *, ids=None, table=None, emb=None):
        src = ids if ids is not None else emb
        self.B, self.S, self.L = src.shape[0], src.shape[1], src.shape[2]
        dev = table.device if table is not None else emb.device
        if emb is not None and emb.requires_grad:
            raise RuntimeError("umpr_b200: the embedding is frozen on this path (model.py:237); no input gradient is produced")
        # model.py:42-43 flatten + model.py:18.  A data-pipeline thread may have prepared the plan already (train.PlanPrefetcher
        # attaches it to the lengths tensor); it is only accepted if it was built for exactly this tensor and padded length.
        pre = getattr(lengths, "_umpr_plan", None)
        if pre is not None and pre.L == self.L and pre.N == lengths.numel() and pre.device == torch.device(dev):
            self.plan = pre.ensure_uploaded()
        else:
            self.plan = PackPlan(lengths.reshape(-1), self.L, dev)
        self._src = (dict(table=table, ids=ids.reshape(self.B * self.S, self.L)) if ids is not None
                     else dict(dense=emb.reshape(self.B * self.S, self.L, emb.shape[-1])))
        self.E = table.shape[1] if ids is not None else emb.shape[-1]
        self._xp = None
        # plans with 128-row tiles feed the fused tensor-core GRU: tokens packed straight into bf16 hi/lo operand images;
        # the fp32 packing (CUDA-core kernels of small sides) is only materialised if somebody asks for it
        self.xq = F.gather_pack_tc(self.plan, **self._src)[0] if (self.plan.R == 128 and F.TENSOR_CORE_GRU) else None

    @property
    def xp(self):
        if self._xp is None:
            self._xp = F.gather_pack(self.plan, **self._src)[0]
        return self._xp


def _tag(gru_repr, plan, sink=None):
    """Attach the pack plan that produced an ImprovedRnn output (rows beyond each sentence's length are exactly zero) - and the
    version counter of the tensor at that moment: an in-place modification afterwards invalidates the tag (``_tagged_plan``)."""
    gru_repr._umpr_plan, gru_repr._umpr_sink, gru_repr._umpr_version = plan, sink, gru_repr._version


def _tagged_plan(gru_repr):
    """→ (plan, sink) if ``gru_repr`` is an untouched ImprovedRnn output of this library, else (None, None)."""
    if getattr(gru_repr, "_umpr_plan", None) is None or getattr(gru_repr, "_umpr_version", -1) != gru_repr._version:
        return None, None
    return gru_repr._umpr_plan, getattr(gru_repr, "_umpr_sink", None)


def _gru_weights(gru: nn.GRU):
    return [gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0,
            gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse]


class ImprovedRnn(nn.Module):
    """model.py:6-21.  ``forward(data (N,L,E), lengths (N,)) -> (result (N,L,128), hidden (2,N,64))`` with
    ``result[n] = GRU(data[unsorted_indices[n]])`` zero-padded to ``total_length = data.shape[1]``."""

    def __init__(self, module, *args, **kwargs):
        assert module in (nn.RNN, nn.LSTM, nn.GRU)
        super().__init__()
        self.module = module(*args, **kwargs)
        m = self.module
        if not (isinstance(m, nn.GRU) and m.batch_first and m.bidirectional and m.num_layers == 1 and m.hidden_size == F.H
                and m.bias and m.input_size < F.KP):
            raise NotImplementedError("umpr_b200 builds the reference's configuration only: nn.GRU, batch_first, "
                                      "bidirectional, 1 layer, hidden_size=64, input_size<64")

    def forward(self, data, lengths):
        with _on(self.module.weight_ih_l0.device):
            if isinstance(data, PackedReviews):
                pk = data
            else:
                pk = PackedReviews(lengths, emb=data.unsqueeze(0))
            return self.run(pk)

    def run(self, pk: PackedReviews, want_hidden=True):
        return F.gru_forward(pk.plan, pk.xp if pk.xq is None else None, pk.E, _gru_weights(self.module), want_hidden, xq=pk.xq)

    def run_many(self, pks, want_hidden=False):
        """Several review sides through this GRU (the reference calls it once per side with the same weights)."""
        return F.gru_forward_multi([pk.plan for pk in pks], [pk.xp if pk.xq is None else None for pk in pks], [pk.xq for pk in pks], pks[0].E,
                                   _gru_weights(self.module), want_hidden)


class RNet(nn.Module):
    """model.py:24-56."""

    def __init__(self, gru_in, gru_out, pretrained: str = None):
        super().__init__()
        self.gru = ImprovedRnn(nn.GRU, input_size=gru_in, hidden_size=gru_out, batch_first=True, bidirectional=True)
        self.M = nn.Parameter(torch.randn(2 * gru_out, 2 * gru_out))
        if pretrained is not None:
            _load_pretrained(self, pretrained, "R-Net")

    def forward(self, user_emb, item_emb, u_lengths, i_lengths):
        with _on(self.M.device):
            pu = user_emb if isinstance(user_emb, PackedReviews) else PackedReviews(u_lengths, emb=user_emb)
            pi = item_emb if isinstance(item_emb, PackedReviews) else PackedReviews(i_lengths, emb=item_emb)
            (gru_u, _), (gru_i, _) = self.gru.run_many([pu, pi])           # model.py:45-46: shared weights, one fused launch
            gru_u = gru_u.view(pu.B, pu.S * pu.L, -1)
            gru_i = gru_i.view(pi.B, pi.S * pi.L, -1)
            sinks = (F.GradSink(), F.GradSink())
            # valid-row tables pay for their host-side cost on large sides only (the small-batch regime is host-bound)
            plans = (pu.plan, pi.plan) if (pu.plan.R == 128 and pi.plan.R == 128) else None
            soft_u, soft_i, atte_u, atte_i = F.co_attention(gru_u, gru_i, self.M, plans=plans, sinks=sinks)
            _tag(gru_u, pu.plan, sinks[0])          # lets S-Net skip the positions beyond each sentence's length
            _tag(gru_i, pi.plan, sinks[1])          # and hand its dx to the co-attention backward (F.GradSink)
            return gru_u, gru_i, soft_u, soft_i, atte_u, atte_i


class SNet(nn.Module):
    """model.py:59-81."""

    def __init__(self, self_atte_size, repr_size, pretrained: str = None):
        super().__init__()
        self.Ms = nn.Parameter(torch.randn(self_atte_size, repr_size))
        self.Ws = nn.Parameter(torch.randn(1, self_atte_size))
        if pretrained is not None:
            _load_pretrained(self, pretrained, "S-Net")

    def forward(self, gru_repr, word_soft, sent_length):
        plan, sink = _tagged_plan(gru_repr)
        with _on(self.Ms.device):
            return F.s_net(gru_repr, word_soft, sent_length, self.Ms, self.Ws, plan=plan, sink=sink)


class CNet(nn.Module):
    """model.py:84-126."""

    def __init__(self, gru_in, gru_out, k_count, k_size, view_size, threshold=0.35, pretrained: str = None):
        super().__init__()
        self.threshold = threshold
        self.gru = ImprovedRnn(nn.GRU, input_size=gru_in, hidden_size=gru_out, batch_first=True, bidirectional=True)
        self.cnn = nn.Sequential(
            nn.Conv1d(in_channels=2 * gru_out, out_channels=k_count, kernel_size=k_size, padding=(k_size - 1) // 2),
            nn.ReLU(),
        )
        self.linear = nn.Sequential(nn.Linear(k_count, view_size), nn.Sigmoid())
        if pretrained is not None:
            _load_pretrained(self, pretrained, "S-Net")

    def forward(self, review_emb, lengths):
        with _on(self.cnn[0].weight.device):
            pk = review_emb if isinstance(review_emb, PackedReviews) else PackedReviews(lengths, emb=review_emb)
            return self.forward_many([pk])[0]

    def forward_many(self, pks):
        """``forward`` for several review sides (model.py:182-184 calls C-Net on ui, user and item with shared weights):
        their GRUs run as one fused launch, the convolution tails per side."""
        res = []
        for pk, (gru_repr, _) in zip(pks, self.gru.run_many(pks)):
            gru_repr = gru_repr.view(pk.B, pk.S * pk.L, -1)
            _tag(gru_repr, pk.plan)
            view_p, final_repr = F.c_net_tail(gru_repr, pk.S, pk.L, self.cnn[0].weight, self.cnn[0].bias,
                                              self.linear[0].weight, self.linear[0].bias, self.threshold, plan=pk.plan if pk.plan.R == 128 else None)
            res.append((gru_repr, view_p, final_repr))
        return res


class SSNet(nn.Module):
    """model.py:129-143.  Inside ``ControlNet`` it is fused with Eq.18; the standalone forward (trainable) is its own small kernel."""

    def __init__(self, input_size, pretrained: str = None):
        super().__init__()
        self.linear = nn.Sequential(nn.Linear(input_size, 1), nn.Sigmoid())
        if pretrained is not None:
            _load_pretrained(self, pretrained, "SS-Net")

    def forward(self, sentiment_emb):
        lin = self.linear[0]
        with _on(lin.weight.device):
            return F.ss_net(sentiment_emb, lin.weight, lin.bias)


class ReviewNet(nn.Module):
    """model.py:146-169."""

    def __init__(self, emb_size, gru_size, atte_size):
        super().__init__()
        self.r_net = RNet(emb_size, gru_size)
        self.s_net_u = SNet(atte_size, gru_size * 2)
        self.s_net_i = SNet(atte_size, gru_size * 2)
        self.linear_u = nn.Linear(gru_size * 4, gru_size * 2, bias=False)
        self.linear_i = nn.Linear(gru_size * 4, gru_size * 2, bias=False)

    def forward(self, user_emb, item_emb, u_lengths, i_lengths):
        u_s_length = user_emb.L if isinstance(user_emb, PackedReviews) else user_emb.shape[-2]
        i_s_length = item_emb.L if isinstance(item_emb, PackedReviews) else item_emb.shape[-2]
        with _on(self.linear_u.weight.device):
            gru_u, gru_i, soft_u, soft_i, atte_u, atte_i = self.r_net(user_emb, item_emb, u_lengths, i_lengths)
            _, sentiment_u = self.s_net_u(gru_u, soft_u, u_s_length)
            _, sentiment_i = self.s_net_i(gru_i, soft_i, i_s_length)
            return F.text_match(atte_u, sentiment_u, atte_i, sentiment_i, self.linear_u.weight, self.linear_i.weight)


class ControlNet(nn.Module):
    """model.py:172-198."""

    def __init__(self, emb_size, gru_size, k_count, k_size, view_size, threshold, atte_size):
        super().__init__()
        self.c_net = CNet(emb_size, gru_size, k_count, k_size, view_size, threshold)
        self.s_net = SNet(atte_size, repr_size=gru_size * 2)
        self.ss_net = SSNet(input_size=gru_size * 2)

    def forward(self, user_emb, item_emb, ui_emb, u_lengths, i_lengths, ui_lengths):
        ui_s_length = ui_emb.L if isinstance(ui_emb, PackedReviews) else ui_emb.shape[-2]
        lin = self.ss_net.linear[0]
        with _on(lin.weight.device):
            pks = [e if isinstance(e, PackedReviews) else PackedReviews(l, emb=e)
                   for e, l in ((ui_emb, ui_lengths), (user_emb, u_lengths), (item_emb, i_lengths))]
            (gru_repr, view_p, c_net_out), (_, _, c_u), (_, _, c_i) = self.c_net.forward_many(pks)      # model.py:182-184
            s, _ = self.s_net(gru_repr, view_p, ui_s_length)
            prefer_pos, prefer_neg = F.control_tail(s, view_p, c_net_out, lin.weight, lin.bias, EQ18_EPS)
            return c_u, c_i, prefer_pos, prefer_neg


class VisualNet(nn.Module):
    """model.py:201-229 without the VGG16 backbone (out of scope): ``images`` are the backbone's
    output features ``(B, V, Pc, vgg_out[, 1, 1])``, i.e. what ``self.vgg16(images).view(...)`` yields at model.py:218."""

    def __init__(self, view_size, vgg_out=1000):
        super().__init__()
        self.pos_v_emb = nn.Parameter(torch.randn(view_size, vgg_out))
        self.neg_v_emb = nn.Parameter(torch.randn(view_size, vgg_out))
        self.linear = nn.Linear(vgg_out, 1)

    def forward(self, images, c_u, c_i):
        feat = images.reshape(images.shape[0], images.shape[1], images.shape[2], -1)
        with _on(self.pos_v_emb.device):
            return F.visual_tail(feat, c_u, c_i, self.pos_v_emb, self.neg_v_emb, self.linear.weight, self.linear.bias)


class UMPR(nn.Module):
    """model.py:232-278.  ``forward`` takes the 8-tuple of ``dataset.py:173-182`` and returns ``(prediction, loss)``."""

    def __init__(self, config, word_emb):
        super().__init__()
        self.review_net_only = config.review_net_only
        self.loss_v_rate = config.loss_v_rate
        self.embedding = nn.Embedding.from_pretrained(torch.as_tensor(word_emb, dtype=torch.float32).clone())
        self.review_net = ReviewNet(self.embedding.embedding_dim, config.gru_size, config.self_atte_size)
        if config.review_net_only:
            self.linear_fusion = nn.Sequential(nn.Linear(config.gru_size * 2, 1), nn.ReLU())
        else:
            view_size = len(config.views)
            self.control_net = ControlNet(self.embedding.embedding_dim, config.gru_size, config.kernel_count,
                                          config.kernel_size, view_size, config.threshold, config.self_atte_size)
            self.visual_net = VisualNet(view_size)
            self.linear_fusion = nn.Sequential(nn.Linear(config.gru_size * 2 + view_size + view_size, 1), nn.ReLU())

    def forward(self, user_reviews, item_reviews, ui_reviews, u_lengths, i_lengths, ui_lengths, photos, labels):
        table = self.embedding.weight
        device = table.device
        if device.type != "cuda":
            raise RuntimeError("umpr_b200: move the model to a CUDA device first (there is no CPU path)")
        with _on(device):
            return self._forward(table, device, user_reviews, item_reviews, ui_reviews, u_lengths, i_lengths, ui_lengths, photos, labels)

    def _native_eval(self):
        """``umpr_step`` bound to this model for forward-only use (evaluate.py:8-11); rebuilt when the parameters have moved."""
        from .step import NativeStep
        st = getattr(self, "_eval_step", None)
        if st is not None:
            try:
                st.check_addresses()
            except RuntimeError:
                st = None
        if st is None:
            st = NativeStep(self, with_grads=False)
            object.__setattr__(self, "_eval_step", st)
        return st

    def _forward(self, table, device, user_reviews, item_reviews, ui_reviews, u_lengths, i_lengths, ui_lengths, photos, labels):
        if not torch.is_grad_enabled() and NATIVE_EVAL:
            # no autograd tape wanted: the whole forward as one native call when the batch lies inside that path's envelope
            st = self._native_eval()
            batch = (user_reviews, item_reviews, ui_reviews, u_lengths, i_lengths, ui_lengths, photos, labels)
            plans = st.plans_of(batch, device)
            if st.supported(batch, plans):
                return st.run(batch, False, plans, routing_log=F.ROUTING_LOG)
        to = lambda v: v.to(device, non_blocking=True)
        user_reviews, item_reviews = to(user_reviews), to(item_reviews)
        labels = to(labels)
        pu = PackedReviews(u_lengths, ids=user_reviews, table=table)          # model.py:262-263 fused into the pack
        pi = PackedReviews(i_lengths, ids=item_reviews, table=table)
        review_net_repr = self.review_net(pu, pi, u_lengths, i_lengths)
        lf = self.linear_fusion[0]
        if self.review_net_only:
            prediction = F.fusion(review_net_repr, None, None, lf.weight, lf.bias)
            loss = F.umpr_loss(prediction, labels)
        else:
            photos = to(photos)
            pui = PackedReviews(ui_lengths, ids=to(ui_reviews), table=table)
            c_u, c_i, prefer_pos, prefer_neg = self.control_net(pu, pi, pui, u_lengths, i_lengths, ui_lengths)
            pos_match, neg_match, final_pos, final_neg = self.visual_net(photos, c_u, c_i)
            prediction = F.fusion(review_net_repr, final_pos, final_neg, lf.weight, lf.bias)
            loss = F.umpr_loss(prediction, labels, prefer_pos, prefer_neg, pos_match, neg_match, self.loss_v_rate)
        return prediction, loss
