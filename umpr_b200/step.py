"""The whole UMPR step as one C-ABI call (``umpr_step``, csrc/step.cu): model.py:257-278 forward and - in training - its backward,
issued from native host code out of one workspace.  The Python autograd path (``functional.py``) stays the general form (any module
on its own, any shape); this is the fast form of the standard model on the tensor-core path, used by ``train.FlatTrainer`` and by
``UMPR.forward`` under ``no_grad``.

What remains on the host per step: the reference's own ``torch.sort`` call and the integer plans derived from it (``plan.py``,
prepared on a worker thread by ``train.PlanPrefetcher``), one pinned upload per review side, two tile schedules, one ctypes call.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .plan import PackPlan, build_schedule, upload_int32

MAX_VALID = 2560         # valid positions per sample the tensor-core co-attention handles (csrc/coattn_tc.cu): S=20 sentences of L=128


def _gru(gru):
    return [gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0,
            gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse]


class NativeStep:
    """Binds a ``umpr_b200.UMPR`` to ``umpr_step``.  Parameter (and gradient) addresses are read at construction - build it after the
    parameters have reached their final storage (``FlatTrainer`` makes them views of its flat buffers) and rebuild it if they move."""

    def __init__(self, model, with_grads: bool):
        rn, full = model.review_net, not model.review_net_only
        self.model, self.full, self.with_grads = model, full, with_grads
        lf = model.linear_fusion[0]
        tensors = {"rnet_gru": _gru(rn.r_net.gru.module), "M": rn.r_net.M, "snet_u_Ms": rn.s_net_u.Ms, "snet_u_Ws": rn.s_net_u.Ws,
                   "snet_i_Ms": rn.s_net_i.Ms, "snet_i_Ws": rn.s_net_i.Ws, "lin_u": rn.linear_u.weight, "lin_i": rn.linear_i.weight,
                   "fus_w": lf.weight, "fus_b": lf.bias}
        V = Pc = F = KC = 0
        ks = 3
        if full:
            cn, vn = model.control_net, model.visual_net
            conv, clin = cn.c_net.cnn[0], cn.c_net.linear[0]
            tensors.update({"cnet_gru": _gru(cn.c_net.gru.module), "conv_w": conv.weight, "conv_b": conv.bias, "clin_w": clin.weight,
                            "clin_b": clin.bias, "csnet_Ms": cn.s_net.Ms, "csnet_Ws": cn.s_net.Ws, "ss_w": cn.ss_net.linear[0].weight,
                            "ss_b": cn.ss_net.linear[0].bias, "pos_e": vn.pos_v_emb, "neg_e": vn.neg_v_emb, "vis_w": vn.linear.weight,
                            "vis_b": vn.linear.bias})
            V, F = vn.pos_v_emb.shape
            KC, _, ks = conv.weight.shape
            self.threshold = float(cn.c_net.threshold)
        table = model.embedding.weight
        self.device = table.device
        m = _lib.StepModel()
        m.review_net_only, m.V, m.Pc, m.F, m.KC, m.ksize, m.E = int(not full), V, 1, F, KC, ks, table.shape[1]
        m.threshold = self.threshold if full else 0.0
        m.loss_v_rate = float(model.loss_v_rate)
        from .model import EQ18_EPS
        m.eq18_eps = EQ18_EPS
        m.table = table.data_ptr()
        self._keep = [table]
        for name, t in tensors.items():
            ts = t if isinstance(t, list) else [t]
            for x in ts:
                if x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
                    raise RuntimeError(f"umpr_b200: parameter {name} must be contiguous fp32 on {self.device}")
                if with_grads and (x.grad is None or not x.grad.is_contiguous() or x.grad.dtype != torch.float32):
                    raise RuntimeError(f"umpr_b200: parameter {name} has no fp32 gradient buffer to accumulate into")
            self._keep += ts
            if isinstance(t, list):
                setattr(m, name, (C.c_void_p * 8)(*[x.data_ptr() for x in ts]))
                if with_grads:
                    setattr(m, "g_" + name, (C.c_void_p * 8)(*[x.grad.data_ptr() for x in ts]))
            else:
                setattr(m, name, t.data_ptr())
                if with_grads:
                    setattr(m, "g_" + name, t.grad.data_ptr())
        self.m = m
        self._addr = [(x, x.data_ptr(), x.grad.data_ptr() if with_grads else 0) for x in self._keep[1:]]
        self._ws = None
        self._zero = torch.zeros(2 * 128 * 128, dtype=torch.uint8, device=self.device)

    def check_addresses(self):
        """Parameters or gradient buffers that moved since construction (``model.to``, ``zero_grad(set_to_none=True)``, another
        optimizer) would make the kernels write into freed memory: refuse loudly."""
        for x, pa, ga in self._addr:
            if x.data_ptr() != pa or (self.with_grads and (x.grad is None or x.grad.data_ptr() != ga)):
                raise RuntimeError("umpr_b200: a parameter or its gradient buffer moved after NativeStep was built; rebuild it")

    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def plans_of(batch, device):
        """The three PackPlans of a batch (prepared ahead by ``train.prepare_batch`` when they ride on the lengths tensors)."""
        out = []
        for ids, lens in ((batch[0], batch[3]), (batch[1], batch[4]), (batch[2], batch[5])):
            if ids.dim() != 3 or lens.numel() == 0:
                out.append(None)
                continue
            pre = getattr(lens, "_umpr_plan", None)
            if pre is not None and pre.L == ids.shape[2] and pre.N == lens.numel() and pre.device == torch.device(device):
                out.append(pre)
            else:
                out.append(PackPlan(lens.reshape(-1), ids.shape[2], device, upload=False))
        return out

    def supported(self, batch, plans) -> bool:
        """The envelope of ``umpr_step``: tensor-core tiles on every side, sentence lengths the tiled kernels take, at most
        ``MAX_VALID`` valid positions per sample."""
        n = 3 if self.full else 2
        if batch[0].shape[:2] != batch[1].shape[:2] or batch[0].shape[2] != batch[1].shape[2]:
            return False
        for k in range(n):
            p = plans[k]
            if p is None or p.R != 128 or p.L > (126 if self.full else 128):
                return False
            if k < 2 and p.max_valid_per_sample(batch[k].shape[0]) > MAX_VALID:
                return False
        return True

    def work_table(self, batch, plans):
        """Algorithmic work of one training step per entry point, {name: [flops, bytes]} (SURVEY.md §8d / DESIGN.md §4 formulas over the
        VALID tokens of each side) - what bench.py divides by the measured kernel time."""
        H, D, ATT, KP = 64, 128, 64, 64
        E, KC = self.m.E, self.m.KC
        B = batch[0].shape[0]
        tok = [float(p.tokens) if p is not None else 0.0 for p in plans]
        nsent = [float(p.N) if p is not None else 0.0 for p in plans]
        n = 3 if self.full else 2
        launches = [tok[0] + tok[1]] + ([tok[0] + tok[1] + tok[2]] if self.full else [])
        P = batch[0].shape[1] * batch[0].shape[2]
        w = {"umpr_gather_pack_tc": [0.0, sum(tok[:n]) * (8 + 4 * E + 4 * KP)],
             "umpr_gru_fwd_tc": [sum(2.0 * t * (E + H) * 6 * H for t in launches), 0.0],
             # recurrence backward (dh and dW_hh) 98 304 + dW_ih 38 400 FLOP per token; the recomputed gates are not counted
             "umpr_gru_bwd_tc": [sum(2.0 * t * 2 * H * 3 * H + 2.0 * t * 2 * 3 * H * (E + H) for t in launches), 0.0],
             "umpr_tc_gemm_ws": [2 * 2.0 * tok[1] * D * D, 0.0],
             "umpr_tc_gemm_tn": [2.0 * D * D * B * P, 0.0],
             "umpr_coattn_fwd_tc": [2.0 * tok[0] * tok[1] / B * D, 2.0 * (tok[0] + tok[1]) * D * 4],
             "umpr_coattn_bwd": [0.0, 3.0 * (tok[0] + tok[1]) * D * 4],
             "umpr_snet_fwd_tc": [sum(2.0 * t * D * ATT for t in tok[:n]), sum(t * D * 8.0 for t in tok[:n])],
             "umpr_snet_bwd_tc": [sum(6.0 * t * D * ATT for t in tok[:n]), sum(t * D * 8.0 for t in tok[:n])]}
        if self.full:
            w["umpr_cnet_conv_fwd_tc"] = [sum(2.0 * (t + s) * 3 * D * KC for t, s in zip(tok, nsent)), sum(t * D * 4.0 for t in tok)]
            for k in ("umpr_cnet_conv_bwd_dx_tc", "umpr_cnet_conv_bwd_dw_tc"):
                w[k] = [sum(2.0 * (t + 2 * s) * 3 * D * KC for t, s in zip(tok, nsent)), sum(t * D * 4.0 + s * KC * 8.0 for t, s in zip(tok, nsent))]
        return w

    @staticmethod
    def profile_begin(only=None):
        """Time the entry points inside the following ``run`` calls (all, or only the named one) with CUDA events on the stream."""
        if _lib.load().umpr_step_profile_begin(only.encode() if only else None) != 0:
            raise RuntimeError(f"umpr_step_profile_begin: {_lib.last_error()}")

    @staticmethod
    def serialize(on: bool):
        """Issue the whole step on the caller's stream (``on``) instead of the library's side streams: per-kernel CUDA-event times
        are then free of co-scheduled kernels."""
        if _lib.load().umpr_step_streams(int(bool(on))) != 0:
            raise RuntimeError(f"umpr_step_streams: {_lib.last_error()}")

    @staticmethod
    def profile_end():
        """→ {entry point: dict(calls, ms)}; synchronises."""
        n_max = 64
        names = C.create_string_buffer(48 * n_max)
        ms, calls, n = (C.c_float * n_max)(), (C.c_int * n_max)(), C.c_int(0)
        if _lib.load().umpr_step_profile_end(n_max, names, ms, calls, C.byref(n)) != 0:
            raise RuntimeError(f"umpr_step_profile_end: {_lib.last_error()}")
        return {names.raw[48 * i:48 * (i + 1)].split(b"\0", 1)[0].decode(): dict(calls=calls[i], ms=ms[i]) for i in range(n.value)}

    def run(self, batch, train: bool, plans=None, routing_log=None):
        """→ (prediction (B,), loss scalar).  ``train``: also the backward, parameter gradients accumulated into ``p.grad``."""
        if train and not self.with_grads:
            raise RuntimeError("umpr_b200: this NativeStep was built without gradient buffers")
        dev = self.device
        n = 3 if self.full else 2
        plans = plans or self.plans_of(batch, dev)
        B = batch[0].shape[0]
        sides = (_lib.StepSide * 3)()
        keep = []
        for k in range(n):
            p = plans[k].ensure_uploaded()
            ids = batch[k]
            ids = ids.to(dev, non_blocking=True) if ids.device != dev else ids
            ids = ids.contiguous()
            st, s_nt = p.snet_table()
            ct, c_nt = p.cnet_table() if self.full else (None, 0)
            keep += [ids, st, ct, p]
            sides[k] = _lib.StepSide(ids.data_ptr(), p.buf.data_ptr(), st.data_ptr(), ct.data_ptr() if ct is not None else None, s_nt, c_nt,
                                     p.n_tiles, p.n_slabs, B, ids.shape[1], ids.shape[2], p.max_valid_per_sample(B))
        from . import plan as plan_mod
        n_ctas = max(1, _lib.sm_count(dev) // 2)
        n_ctas_r = max(1, (_lib.sm_count(dev) - plan_mod.SMS_RESERVED_FOR_COMM) // 2)
        pre = getattr(batch[3], "_umpr_sched", None) if plans[0] is getattr(batch[3], "_umpr_plan", None) else None     # built by the prefetch worker
        sr, nq_r = pre[0] if pre else build_schedule([plans[0].tile_len, plans[1].tile_len], n_ctas_r)
        sr = upload_int32(sr, dev)
        sc, nq_c = None, 0
        photos = None
        if self.full:
            sc, nq_c = pre[1] if pre and len(pre) > 1 else build_schedule([plans[2].tile_len, plans[0].tile_len, plans[1].tile_len], n_ctas)
            sc = upload_int32(sc, dev)
            photos = batch[6].to(dev, non_blocking=True)
            photos = photos.reshape(B, photos.shape[1], photos.shape[2], -1).to(torch.float32).contiguous()
            if photos.shape[1] != self.m.V or photos.shape[3] != self.m.F:
                raise RuntimeError(f"umpr_b200: photos must be (B, {self.m.V}, photo_count, {self.m.F}) features, got {tuple(photos.shape)}")
            self.m.Pc = photos.shape[2]
        labels = batch[7].to(dev, non_blocking=True).to(torch.float32).contiguous()
        m = self.m
        taps = None
        if routing_log is not None:
            P = batch[0].shape[1] * batch[0].shape[2]
            taps = [torch.empty(2, B, P, dtype=torch.int32, device=dev)]
            m.routing_coattn = taps[0].data_ptr()
            if self.full:
                for j, k in enumerate((2, 0, 1)):       # the order control_net calls c_net in: ui, user, item
                    t = torch.empty(B * batch[k].shape[1], self.m.KC, dtype=torch.int32, device=dev)
                    taps.append(t)
                    m.routing_cnet[k] = t.data_ptr()
        lib = _lib.load()
        need = C.c_longlong(0)
        if lib.umpr_step_workspace_bytes(C.byref(m), sides, int(train), C.byref(need)) != 0:
            raise RuntimeError(f"umpr_step_workspace_bytes: {_lib.last_error()}")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None                                    # release before growing
            self._ws = torch.empty(int(need.value * 1.25) + 4096, dtype=torch.uint8, device=dev)
        pred = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.call("umpr_step", C.addressof(m), C.addressof(sides), _lib.ptr(photos), _lib.ptr(labels), _lib.ptr(sr), nq_r, _lib.ptr(sc), nq_c,
                  _lib.ptr(self._zero), _lib.ptr(self._ws), self._ws.numel(), _lib.ptr(pred), _lib.ptr(loss), int(train))
        # kernels launched inside the call (csrc/step.cu), for bench.py's launch count: the call itself was counted as one
        fwd = (1 + 1 + 1 + 3 + 4 + 1 + 1 + 1) + ((1 + 10 + 1 + 1 + 2) if self.full else 0)      # one gather launch; conv: 1 prep + 3 x (conv, fix, head)
        bwd = ((1 + 1 + 1 + 1 + 1 + 1 + 6 + 1 + 1 + 1 + 1) + ((3 + 1 + 1 + 12 + 1 + 1) if self.full else 0)) if train else 0
        _lib.launch_count += fwd + bwd - 1
        if taps is not None:
            routing_log.append(("coattn", taps[0]))
            for t in taps[1:]:
                routing_log.append(("cnet", t))
            m.routing_coattn = None
            for k in range(3):
                m.routing_cnet[k] = None
        return pred, loss
