"""autograd Functions over the C-ABI kernels.  PyTorch here only owns memory, streams and the autograd tape.

Each Function cites the reference lines (``src/model.py``) whose forward+backward it replaces.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import call, ptr, ptr_array
from .plan import PackPlan

H, D, ATT, KP, SV = 64, 128, 64, 64, 256
TENSOR_CORE_GRU = True         # fused tcgen05 input projection + recurrence (gru_rec_tc.cu) for plans with 128-row tiles
TENSOR_CORE_WGRAD = True       # tcgen05 GRU weight gradients with MN-major operands (gru_wgrad_tc.cu)
TENSOR_CORE_CONV = True        # tcgen05 implicit-GEMM convolution (cnet_tc.cu)
TENSOR_CORE_CONV_DX = True     # conv input gradient on tcgen05 (cnet_bwd_tc.cu) when the input carries its pack plan
TENSOR_CORE_SNET = True        # tcgen05 S-Net over the valid positions only (snet_tc.cu), for inputs that carry their pack plan
TENSOR_CORE_GEMM = True        # tcgen05 3xBF16 dense products (gemm_tc.cu); False = fp32 CUDA-core kernel everywhere
TENSOR_CORE_COATTN = os.environ.get("UMPR_TC_COATTN", "1") == "1"     # tcgen05 affinity (coattn_tc.cu) is correct but epilogue-bound (profiles/r1b notes); fp32 kernel is faster for now


GRU_SCHED_CTAS = None          # tests: CTAs per direction the fused GRU launches are scheduled onto (default: half the SMs)
ROUTING_LOG = None             # tests: a list that receives ("coattn", arg (2,B,P)) / ("cnet", cidx (N,KC)) - the arg-max positions the kernels chose
POISON_UNWRITTEN = False       # tests: fill the gradient rows the kernels leave unwritten (positions beyond a sentence's length, which the
                               # packed GRU backward never reads, model.py:18) with NaN - parameter gradients must not change
DIRECT_GRAD_ACCUM = False      # set by train.FlatTrainer for the duration of its backward pass (see _sinks)


_WS = {}


def _workspace_floats(entry: str, a: int, b: int = 0) -> int:
    """Scratch size (in fp32 elements) of an entry point that takes a caller-owned workspace, from ``umpr_workspace_bytes``."""
    key = (entry, a, b)
    if key not in _WS:
        import ctypes
        out = ctypes.c_longlong(0)
        if _lib.load().umpr_workspace_bytes(entry.encode(), a, b, ctypes.byref(out)) != 0:
            raise RuntimeError(f"umpr_workspace_bytes: {_lib.last_error()}")
        _WS[key] = (out.value + 3) // 4
    return _WS[key]


def _sinks(params):
    """Destinations of parameter gradients that the kernels ACCUMULATE (+=) → (buffers to hand to the kernels, tensors to return
    to autograd).  Under train.FlatTrainer every parameter's ``.grad`` is a view into the flat gradient bucket (zeroed each step), so
    the kernels add straight into it and autograd gets ``None``: no temporary, no zero-fill kernel, no AccumulateGrad add kernel per
    parameter.  Anywhere else (``torch.autograd.grad``, plain ``backward`` on a fresh model) fresh zeroed buffers are returned."""
    if DIRECT_GRAD_ACCUM and all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous() and p.grad.shape == p.shape
                                 for p in params):
        return [p.grad for p in params], [None] * len(params)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    out, o = [], 0
    for p in params:
        out.append(flat[o:o + p.numel()].view(p.shape))
        o += p.numel()
    return out, out


def _grad_like(x):
    """Gradient buffer for a GRU output.  With a pack plan the kernels write the rows below each sentence's length only; the rest is
    never read downstream (the packed GRU backward skips them) and stays unwritten - or NaN under POISON_UNWRITTEN."""
    return torch.full_like(x, float("nan")) if POISON_UNWRITTEN else torch.empty_like(x)


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _chk(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"umpr_b200: {what} must live on a CUDA device (there is no CPU path)")
    return t


def _n_ctas(device, per_sm=1):
    return _lib.sm_count(device) * per_sm


# --------------------------------------------------------------------------------------------------------------------
# embedding gather + pack   (model.py:262-264 and the pack half of model.py:18)
# --------------------------------------------------------------------------------------------------------------------
def gather_pack(plan: PackPlan, *, table=None, ids=None, dense=None):
    """Packed GRU input ``xp[n_slabs][R][64]``; either ``table[ids]`` (ids: (Nseq, L) int64) or ``dense`` (Nseq, L, E)."""
    if dense is not None:
        dense = _f32(_chk(dense, "GRU input"))
        E = dense.shape[-1]
        dev = dense.device
    else:
        table = _f32(_chk(table, "embedding table"))
        ids = _chk(ids, "token ids").contiguous()
        assert ids.dtype == torch.int64
        E = table.shape[1]
        dev = table.device
    xp = torch.empty(plan.n_slabs * plan.R * KP, dtype=torch.float32, device=dev)
    call("umpr_gather_pack", ptr(table), ptr(ids), ptr(dense), ptr(plan.buf), plan.n_tiles, plan.n_slabs, plan.R, plan.L, E, ptr(xp),
         work=(0.0, plan.tokens * (8 + 4 * E + 4 * KP)))
    return xp, E


def gather_pack_tc(plan: PackPlan, *, table=None, ids=None, dense=None):
    """Packed GRU input as tensor-core operand images ``xq[n_slabs][hi|lo][128][64 bf16]`` (plans with 128-row tiles)."""
    if plan.R != 128:
        raise RuntimeError("umpr_b200: operand images need a pack plan with 128-row tiles")
    if dense is not None:
        dense = _f32(_chk(dense, "GRU input"))
        E, dev = dense.shape[-1], dense.device
    else:
        table = _f32(_chk(table, "embedding table"))
        ids = _chk(ids, "token ids").contiguous()
        assert ids.dtype == torch.int64
        E, dev = table.shape[1], table.device
    xq = torch.empty(plan.n_slabs * 2 * 128 * 128, dtype=torch.uint8, device=dev)
    call("umpr_gather_pack_tc", ptr(table), ptr(ids), ptr(dense), ptr(plan.buf), plan.n_tiles, plan.n_slabs, plan.L, E, ptr(xq),
         work=(0.0, plan.tokens * (8 + 4 * E + 4 * KP)))
    return xq, E


# --------------------------------------------------------------------------------------------------------------------
# ImprovedRnn   (model.py:12-21)
# --------------------------------------------------------------------------------------------------------------------
class _GruFn(Function):
    @staticmethod
    def forward(ctx, plan: PackPlan, xp, E, want_hidden, *w):
        w = [_f32(t) for t in w]
        dev = xp.device
        N, L, R = plan.N, plan.L, plan.R
        G = torch.empty(plan.n_slabs * 2 * R * 3 * H, dtype=torch.float32, device=dev)
        wp = ptr_array(w)
        call("umpr_gru_inproj_tc", ptr(xp), wp, plan.n_slabs, R, E, ptr(G), _n_ctas(dev), work=(2.0 * plan.tokens * E * 6 * H, 0.0))
        out = torch.empty(N, L, D, dtype=torch.float32, device=dev)
        hn = torch.empty(2, N, H, dtype=torch.float32, device=dev) if want_hidden else None
        need_grad = any(ctx.needs_input_grad[4:])
        sv = torch.empty(plan.n_slabs * 2 * R * SV, dtype=torch.float32, device=dev) if need_grad else None
        call("umpr_gru_recurrence_fwd", ptr(G), wp, ptr(plan.buf), plan.n_tiles, plan.n_slabs, R, N, L, ptr(out), ptr(hn), ptr(sv),
             work=(2.0 * plan.tokens * 2 * H * 3 * H, 0.0))
        del G
        ctx.plan, ctx.E = plan, E
        ctx.save_for_backward(xp, out, sv, *w)
        if hn is None:
            hn = out.new_zeros(0)
            ctx.mark_non_differentiable(hn)
        return out, hn

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out, d_hn):
        plan, E = ctx.plan, ctx.E
        xp, out, sv, *w = ctx.saved_tensors
        dev = out.device
        N, L, R = plan.N, plan.L, plan.R
        d_out = _f32(d_out) if d_out is not None else torch.zeros_like(out)
        d_hn = _f32(d_hn) if (d_hn is not None and d_hn.numel()) else None
        dG = torch.empty(plan.n_slabs * 2 * R * SV, dtype=torch.float32, device=dev)
        wp = ptr_array(w)
        call("umpr_gru_recurrence_bwd", ptr(d_out), ptr(d_hn), ptr(out), ptr(sv), wp, ptr(plan.buf), plan.n_tiles, plan.n_slabs, R,
             N, L, ptr(dG), work=(2.0 * plan.tokens * 2 * H * 3 * H, 0.0))
        flat = torch.zeros(sum(t.numel() for t in w), dtype=torch.float32, device=dev)
        grads, o = [], 0
        for t in w:
            grads.append(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
        call("umpr_gru_wgrad_tc" if TENSOR_CORE_WGRAD else "umpr_gru_wgrad", ptr(dG), ptr(xp), ptr(out), ptr(plan.buf), plan.n_tiles, plan.n_slabs, R, L, E, ptr_array(grads),
             _n_ctas(dev, 2), work=(2.0 * plan.tokens * 2 * 3 * H * (E + H), 0.0))
        return (None, None, None, None, *grads)


def gru_forward(plan: PackPlan, xp, E, weights, want_hidden=True, xq=None):
    """weights: 8 tensors in nn.GRU order (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, then *_reverse)."""
    if TENSOR_CORE_GRU and plan.R == 128 and xq is not None:
        out, hn = gru_forward_multi([plan], [xp], [xq], E, weights, want_hidden)[0]
        return out, hn
    return _GruFn.apply(plan, xp, E, want_hidden, *weights)


class _GruTcFn(Function):
    """Several ImprovedRnn calls that share one nn.GRU (model.py:45-46; model.py:182-184) as ONE fused tensor-core launch:
    input projection + recurrence, W_ih/W_hh resident in shared memory (gru_rec_tc.cu)."""

    @staticmethod
    def forward(ctx, plans, xps, xqs, E, want_hidden, *w):
        ctx.params = w
        w = [_f32(t) for t in w]
        dev = xqs[0].device
        n = len(plans)
        need_grad = any(ctx.needs_input_grad[5:])
        segs = (_lib.GruSeg * n)()
        outs, hns, hqs = [], [], []
        tokens = 0
        for i, (plan, xq) in enumerate(zip(plans, xqs)):
            if plan.R != 128:
                raise RuntimeError("umpr_b200: the tensor-core GRU needs pack plans with 128-row tiles")
            out = torch.empty(plan.N, plan.L, D, dtype=torch.float32, device=dev)
            hn = torch.empty(2, plan.N, H, dtype=torch.float32, device=dev) if want_hidden else None
            # training keeps h_t as operand images (256 B per token and direction); the backward recomputes the gates from them
            hq = torch.empty(plan.n_slabs * 2 * 2 * 128 * 128, dtype=torch.uint8, device=dev) if need_grad else None
            segs[i] = _lib.GruSeg(ptr(xq), ptr(plan.buf), ptr(out), ptr(hn), ptr(hq), plan.n_tiles, plan.n_slabs, plan.N, plan.L)
            outs.append(out); hns.append(hn); hqs.append(hq)
            tokens += plan.tokens
        from .plan import build_schedule, upload_int32
        sched, nq = build_schedule([p.tile_len for p in plans], GRU_SCHED_CTAS or max(1, _n_ctas(dev) // 2))
        sched = upload_int32(sched, dev)
        call("umpr_gru_fwd_tc", C.addressof(segs), n, ptr_array(w), E, ptr(sched), nq,
             work=(2.0 * tokens * (E + H) * 6 * H, 0.0))
        ctx.plans, ctx.E, ctx.n, ctx.sched, ctx.nq = plans, E, n, sched, nq
        ctx.save_for_backward(*xqs, *[t for t in hqs if t is not None], *w)
        ctx.out_shapes = [tuple(o.shape) for o in outs]
        res = list(outs)
        for hn in hns:
            if hn is None:
                hn = outs[0].new_zeros(0)
            res.append(hn)
        if not want_hidden:
            ctx.mark_non_differentiable(*res[n:])
        return tuple(res)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads_in):
        plans, E, n = ctx.plans, ctx.E, ctx.n
        saved = ctx.saved_tensors
        xqs, hqs, w = saved[:n], saved[n:2 * n], list(saved[2 * n:])
        dev = xqs[0].device
        grads, rets = _sinks(ctx.params)
        segs = (_lib.GruBwdSeg * n)()
        keep, tokens = [], 0
        for i, plan in enumerate(plans):
            d_out, d_hn = grads_in[i], grads_in[n + i]
            d_out = _f32(d_out) if d_out is not None else torch.zeros(ctx.out_shapes[i], dtype=torch.float32, device=dev)
            d_hn = _f32(d_hn) if (d_hn is not None and d_hn.numel()) else None
            segs[i] = _lib.GruBwdSeg(ptr(d_out), ptr(d_hn), ptr(xqs[i]), ptr(hqs[i]), ptr(plan.buf), plan.n_tiles,
                                     plan.n_slabs, plan.N, plan.L)
            keep += [d_out, d_hn]
            tokens += plan.tokens
        # one reverse-time launch over every side (same tile queues as the forward): gate recomputation + recurrence + all eight
        # weight gradients.  Algorithmic work: 98 304 FLOP/token for the recurrence backward (dh and dW_hh) +
        # 38 400 for dW_ih; the recomputation of the gates (87 552 FLOP/token) is NOT counted
        call("umpr_gru_bwd_tc", C.addressof(segs), n, ptr_array(w), ptr_array(grads), E, ptr(_zero_image(dev)), ptr(ctx.sched), ctx.nq,
             work=(2.0 * tokens * 2 * H * 3 * H + 2.0 * tokens * 2 * 3 * H * (E + H), 0.0))
        return (None, None, None, None, None, *rets)


_ZERO_IMG = {}


def _zero_image(dev):
    """32 KB of zeros: the h_{t-1} operand image of a sequence's first step (umpr_gru_bwd_tc)."""
    key = torch.device(dev).index
    if key not in _ZERO_IMG:
        _ZERO_IMG[key] = torch.zeros(2 * 128 * 128, dtype=torch.uint8, device=dev)
    return _ZERO_IMG[key]


def gru_forward_multi(plans, xps, xqs, E, weights, want_hidden=False):
    """ImprovedRnn over several review sides with shared GRU weights → [(out, hn), ...] in the order given.
    Sides whose plan has 128-row tiles share one fused tensor-core launch; small sides run the CUDA-core kernels."""
    res = [None] * len(plans)
    tc = [i for i, p in enumerate(plans) if TENSOR_CORE_GRU and p.R == 128 and xqs[i] is not None]
    for i0 in range(0, len(tc), 3):
        grp = tc[i0:i0 + 3]
        r = _GruTcFn.apply(tuple(plans[i] for i in grp), tuple(xps[i] for i in grp), tuple(xqs[i] for i in grp), E, want_hidden, *weights)
        for j, i in enumerate(grp):
            res[i] = (r[j], r[len(grp) + j])
    for i, p in enumerate(plans):
        if res[i] is None:
            res[i] = _GruFn.apply(p, xps[i], E, want_hidden, *weights)
    return res


# --------------------------------------------------------------------------------------------------------------------
# strided GEMM helper
# --------------------------------------------------------------------------------------------------------------------
def sgemm(A, a_strides, B, b_strides, C, ldc, M, N, K, *, splits=1, accumulate=False, bias=None, act=0, rows=None):
    """C = act(acc*C + A·B + bias) with strided operands.  Dense "row-major A times K- or N-contiguous B" products run on
    the tensor cores (tcgen05, 3xBF16 split); split-K reductions and odd strides use the CUDA-core kernel."""
    pa = A if isinstance(A, int) else ptr(A)
    pb = B if isinstance(B, int) else ptr(B)
    pc = C if isinstance(C, int) else ptr(C)
    work = (2.0 * M * N * K, 0.0)
    aligned = not ((pa | pb | pc) & 15) and a_strides[1] == 1 and not (a_strides[0] & 3) and not (ldc & 3)
    if TENSOR_CORE_GEMM and aligned and splits == 1 and act in (0, 1, 2) and M >= 64:
        b_kn = -1
        if b_strides[0] == 1 and not (b_strides[1] & 3):          # B[N][K]
            b_kn, ldb = 0, b_strides[1]
        elif b_strides[1] == 1 and not (b_strides[0] & 3):        # B[K][N]
            b_kn, ldb = 1, b_strides[0]
        if b_kn >= 0:
            if N <= 128 and K <= 128 and not (N & 3) and M >= 1024:   # weights stay in shared memory, A streams
                table, n_row_tiles, L = None, 0, 0
                if rows is not None and rows.N * rows.L == M and rows.L <= 128:
                    # ``rows``: the pack plan behind A - rows beyond each sentence's length are exactly zero and are skipped
                    tab, n_row_tiles = rows.snet_table()
                    table, L = tab.data_ptr(), rows.L
                    work = (2.0 * rows.tokens * N * K, 0.0)
                call("umpr_tc_gemm_ws", pa, a_strides[0], pb, ldb, pc, ldc, M, N, K, int(accumulate), ptr(bias), act, b_kn,
                     table, n_row_tiles, L, _lib.sm_count(torch.cuda.current_device()), work=work)
            else:
                call("umpr_tc_gemm_nt", pa, a_strides[0], pb, ldb, pc, ldc, M, N, K, int(accumulate), ptr(bias), act, b_kn, work=work)
            return
    call("umpr_sgemm", pa, a_strides[0], a_strides[1], pb, b_strides[0], b_strides[1], pc, ldc, M, N, K, splits, int(accumulate),
         ptr(bias), act, work=work)


def _splits_for(K, device):
    return max(1, min(_lib.sm_count(device), K // 256))


# --------------------------------------------------------------------------------------------------------------------
# RNet co-attention   (model.py:50-55)
# --------------------------------------------------------------------------------------------------------------------
class GradSink:
    """Hand-over of one gradient between two autograd Functions that consume the same tensor.  ``gru_u`` feeds the co-attention
    and S-Net, and S-Net's other input is the co-attention's soft-max - so S-Net's backward always runs first.  Instead of returning
    its ``dx`` (autograd would then add it to the co-attention's ``dgu`` with one more full pass over both tensors), S-Net parks it
    here and returns nothing; the co-attention backward folds it into the rows it writes anyway (``add_u`` / ``add_i``).

    The hand-over is only taken when S-Net's ``word_soft`` IS the soft-max this sink's co-attention node produced (``key`` = its
    storage address, checked in ``_SNetTcFn.forward``) - only then is the co-attention backward guaranteed to run after S-Net's and
    to need this gradient.  A parked gradient that is still there when the backward pass ends raises (``_check_consumed``)."""
    __slots__ = ("armed", "dx", "key")

    def __init__(self):
        self.armed, self.dx, self.key = False, None, None

    def park(self, dx):
        self.dx = dx
        from torch.autograd import Variable
        Variable._execution_engine.queue_callback(self._check_consumed)      # runs when the current backward pass has finished

    def _check_consumed(self):
        if self.dx is not None:
            self.dx = None
            raise RuntimeError("umpr_b200: S-Net parked its input gradient for the co-attention backward, which never ran - the "
                               "GRU would silently miss that gradient (call backward on a loss that reaches the co-attention node)")


class _CoAttnFn(Function):
    @staticmethod
    def forward(ctx, plans, sinks, gu, gi, M):
        ctx.sinks = sinks
        if sinks is not None:
            for sk, t in zip(sinks, (gu, gi)):
                sk.armed, sk.dx = bool(t.requires_grad), None
        ctx.params = (M,)
        ctx.cst = (None, 0, 0, None, 0, 0)
        gu, gi, M = _f32(_chk(gu, "gru_u")), _f32(gi), _f32(M)
        B, P, _ = gu.shape
        dev = gu.device
        giM = torch.empty_like(gi)
        ok_plans = plans is not None and all(pl is not None and pl.L <= 128 and pl.N % B == 0 and (pl.N // B) * pl.L == P for pl in plans)
        # the tensor-core kernels handle 2560 positions per sample; with the plans the limit applies to the VALID positions
        pv_max = max(pl.max_valid_per_sample(B) for pl in plans) if ok_plans else P
        use_tc = TENSOR_CORE_COATTN and pv_max <= 2560 and P <= 16384
        use_rows = use_tc and ok_plans
        ctx.rows_i = plans[1] if use_rows else None       # giM / dgi rows beyond each sentence's length are never read: skip them
        sgemm(gi, (D, 1), M, (D, 1), giM, D, B * P, D, D, rows=ctx.rows_i)
        soft = torch.empty(4, B, P, dtype=torch.float32, device=dev)       # soft_u, soft_i, t_u, t_i
        arg = torch.empty(2, B, P, dtype=torch.int32, device=dev)
        atte = torch.empty(2, B, D, dtype=torch.float32, device=dev)
        work = (2.0 * B * P * P * D, 2.0 * B * P * D * 4)
        if use_tc:
            scratch = torch.empty(_workspace_floats("coattn_fwd_tc", B, P), dtype=torch.float32, device=dev)
            cst = [None, None]
            sl = [0, 0, 0, 0]
            if ok_plans:
                # both inputs come out of ImprovedRnn with these plans: rows beyond each sentence's length are exactly zero
                for k, pl in enumerate(plans):
                    table, n_tiles = pl.snet_table()
                    cst[k] = table.data_ptr() + 4 * (n_tiles + 1)
                    sl[2 * k], sl[2 * k + 1] = pl.N // B, pl.L
                ctx.cst = (cst[0], sl[0], sl[1], cst[1], sl[2], sl[3])
                ctx.keep = plans             # keeps the device tables alive until backward
                valid = float(plans[0].tokens) * float(plans[1].tokens) / B
                work = (2.0 * valid * D, 2.0 * (plans[0].tokens + plans[1].tokens) * D * 4)
            call("umpr_coattn_fwd_tc", ptr(gu), ptr(gi), ptr(giM), B, P, cst[0], sl[0], sl[1], cst[1], sl[2], sl[3], pv_max if ok_plans else 0, ptr(scratch),
                 ptr(soft[0]), ptr(soft[1]), ptr(soft[2]), ptr(soft[3]), ptr(arg[0]), ptr(arg[1]), ptr(atte[0]), ptr(atte[1]), work=work)
        else:
            rowkey = torch.empty(B * P, dtype=torch.int64, device=dev)
            colkey = torch.zeros(B * P, dtype=torch.int64, device=dev)
            call("umpr_coattn_fwd", ptr(gu), ptr(gi), ptr(giM), B, P, ptr(rowkey), ptr(colkey), ptr(soft[0]), ptr(soft[1]), ptr(soft[2]),
                 ptr(soft[3]), ptr(arg[0]), ptr(arg[1]), ptr(atte[0]), ptr(atte[1]), work=work)
        if ROUTING_LOG is not None:
            ROUTING_LOG.append(("coattn", arg.clone()))
        ctx.save_for_backward(gu, gi, giM, M, soft, arg)
        if sinks is not None:
            sinks[0].key, sinks[1].key = soft[0].data_ptr(), soft[1].data_ptr()
        return soft[0], soft[1], atte[0], atte[1]

    @staticmethod
    @once_differentiable
    def backward(ctx, d_soft_u, d_soft_i, d_atte_u, d_atte_i):
        gu, gi, giM, M, soft, arg = ctx.saved_tensors
        B, P, _ = gu.shape
        dev = gu.device
        c = lambda t: None if t is None else _f32(t)
        dgu = _grad_like(gu)
        dgi = _grad_like(gi)
        dgiM = torch.empty_like(gi)
        adds = [None, None]
        if ctx.sinks is not None:
            for k, sk in enumerate(ctx.sinks):
                adds[k], sk.dx, sk.armed = sk.dx, None, False
        call("umpr_coattn_bwd", ptr(gu), ptr(gi), ptr(giM), ptr(soft[0]), ptr(soft[1]), ptr(soft[2]), ptr(soft[3]), ptr(arg[0]),
             ptr(arg[1]), ptr(c(d_soft_u)), ptr(c(d_soft_i)), ptr(c(d_atte_u)), ptr(c(d_atte_i)), B, P, *ctx.cst, ptr(adds[0]), ptr(adds[1]), ptr(dgu), ptr(dgi), ptr(dgiM),
             work=(0.0, 6.0 * B * P * D * 4))
        # dgi += dgiM · M^T ;  dM = gi^T · dgiM
        sgemm(dgiM, (D, 1), M, (1, D), dgi, D, B * P, D, D, accumulate=True, rows=ctx.rows_i)
        (dM,), (rM,) = _sinks(ctx.params)
        if B * P >= 4096:      # reduction over every token of the batch: tensor cores, both operands token-major
            call("umpr_tc_gemm_tn", ptr(gi), D, ptr(dgiM), D, ptr(dM), D, D, D, B * P, _n_ctas(dev), work=(2.0 * D * D * B * P, 0.0))
        else:
            sgemm(gi, (1, D), dgiM, (D, 1), dM, D, D, D, B * P, splits=_splits_for(B * P, dev), accumulate=True)
        return None, None, dgu, dgi, rM


def co_attention(gu, gi, M, plans=None, sinks=None):
    """→ soft_u, soft_i (B,P), atte_u, atte_i (B,128).  ``plans`` = (plan_u, plan_i): the PackPlans of the ImprovedRnn calls that
    produced ``gu`` / ``gi`` - their rows beyond each sentence's length are exactly zero and the tensor-core kernels skip them."""
    return _CoAttnFn.apply(plans, sinks, gu, gi, M)


# --------------------------------------------------------------------------------------------------------------------
# SNet   (model.py:71-81)
# --------------------------------------------------------------------------------------------------------------------
class _SNetFn(Function):
    @staticmethod
    def forward(ctx, gru_repr, word_soft, sent_length, Ms, Ws):
        ctx.params = (Ms, Ws)
        x = _f32(_chk(gru_repr, "gru_repr"))
        Ms, Ws = _f32(Ms), _f32(Ws)
        word_soft = _f32(word_soft)
        B = x.shape[0]
        L = int(sent_length)
        S = x.shape[1] // L
        N = B * S
        dev = x.device
        if Ms.shape != (ATT, D) or x.shape[2] != D:
            raise RuntimeError("umpr_b200: SNet is built for self_atte_size=64, repr_size=128")
        train = any(ctx.needs_input_grad)
        self_atte = torch.empty(B, S, D, dtype=torch.float32, device=dev)
        soft = torch.empty(N, L, dtype=torch.float32, device=dev) if train else None
        th = torch.empty(N, L, ATT, dtype=torch.float32, device=dev) if train else None
        call("umpr_snet_fwd", ptr(x), ptr(Ms), ptr(Ws), N, L, ptr(self_atte), ptr(soft), ptr(th), _n_ctas(dev),
             work=(2.0 * N * L * D * ATT, N * L * 4.0 * (D + (ATT if train else 0))))
        Wd = word_soft.numel() // N
        wsum = torch.empty(N, dtype=torch.float32, device=dev)
        sentiment = torch.empty(B, D, dtype=torch.float32, device=dev)
        call("umpr_snet_sentiment_fwd", ptr(self_atte), ptr(word_soft), B, S, Wd, ptr(wsum), ptr(sentiment))
        ctx.dims = (B, S, L, Wd, tuple(word_soft.shape))
        ctx.save_for_backward(x, soft, th, self_atte, wsum, Ms, Ws)
        return self_atte, sentiment

    @staticmethod
    @once_differentiable
    def backward(ctx, d_self_atte, d_sentiment):
        x, soft, th, self_atte, wsum, Ms, Ws = ctx.saved_tensors
        B, S, L, Wd, ws_shape = ctx.dims
        N = B * S
        dev = x.device
        d_sa = torch.empty(N, D, dtype=torch.float32, device=dev)
        want_ws = ctx.needs_input_grad[1] and d_sentiment is not None
        d_wsum = torch.empty(N, dtype=torch.float32, device=dev) if want_ws else None
        call("umpr_snet_sentiment_bwd", ptr(self_atte), ptr(wsum), ptr(None if d_sentiment is None else _f32(d_sentiment)),
             ptr(None if d_self_atte is None else _f32(d_self_atte)), B, S, ptr(d_sa), ptr(d_wsum))
        dx = torch.empty_like(x)
        (dMs, dWs), (rMs, rWs) = _sinks(ctx.params)
        call("umpr_snet_bwd", ptr(x), ptr(th), ptr(soft), ptr(d_sa), ptr(Ms), ptr(Ws), N, L, ptr(dx), ptr(dMs), ptr(dWs), _n_ctas(dev),
             work=(4.0 * N * L * D * ATT, N * L * 4.0 * (2 * D + ATT)))
        d_word_soft = d_wsum.view(N, 1).expand(N, Wd).reshape(ws_shape) if want_ws else None
        return dx, d_word_soft, None, rMs, rWs


class _SNetTcFn(Function):
    """S-Net on the tensor cores for a ``gru_repr`` that came out of ImprovedRnn with ``plan``: only the valid positions are
    multiplied (csrc/snet_tc.cu); the backward recomputes the scores, so nothing but the input is kept."""

    @staticmethod
    def forward(ctx, plan, sink, gru_repr, word_soft, sent_length, Ms, Ws):
        ctx.params = (Ms, Ws)
        # the hand-over is only valid when word_soft is the soft-max of the co-attention node that owns the sink (see GradSink)
        ctx.sink = sink if (sink is not None and sink.key is not None and word_soft.data_ptr() == sink.key) else None
        x = _f32(_chk(gru_repr, "gru_repr"))
        Ms, Ws = _f32(Ms), _f32(Ws)
        word_soft = _f32(word_soft)
        B, L = x.shape[0], int(sent_length)
        S = x.shape[1] // L
        N = B * S
        dev = x.device
        table, n_tiles = plan.snet_table()
        T_v = float(plan.tokens)
        self_atte = torch.empty(B, S, D, dtype=torch.float32, device=dev)
        call("umpr_snet_fwd_tc", ptr(x), ptr(table), n_tiles, ptr(Ms), ptr(Ws), N, L, ptr(self_atte), _n_ctas(dev),
             work=(2.0 * T_v * D * ATT, T_v * D * 8.0))
        Wd = word_soft.numel() // N
        wsum = torch.empty(N, dtype=torch.float32, device=dev)
        sentiment = torch.empty(B, D, dtype=torch.float32, device=dev)
        call("umpr_snet_sentiment_fwd", ptr(self_atte), ptr(word_soft), B, S, Wd, ptr(wsum), ptr(sentiment))
        ctx.dims = (B, S, L, Wd, tuple(word_soft.shape))
        ctx.plan = plan
        ctx.save_for_backward(x, self_atte, wsum, Ms, Ws)
        return self_atte, sentiment

    @staticmethod
    @once_differentiable
    def backward(ctx, d_self_atte, d_sentiment):
        x, self_atte, wsum, Ms, Ws = ctx.saved_tensors
        B, S, L, Wd, ws_shape = ctx.dims
        N = B * S
        dev = x.device
        d_sa = torch.empty(N, D, dtype=torch.float32, device=dev)
        want_ws = ctx.needs_input_grad[3] and d_sentiment is not None
        d_wsum = torch.empty(N, dtype=torch.float32, device=dev) if want_ws else None
        call("umpr_snet_sentiment_bwd", ptr(self_atte), ptr(wsum), ptr(None if d_sentiment is None else _f32(d_sentiment)),
             ptr(None if d_self_atte is None else _f32(d_self_atte)), B, S, ptr(d_sa), ptr(d_wsum))
        dx = _grad_like(x)              # positions beyond a sentence's length stay unwritten: the packed GRU never reads them (model.py:18)
        (dMs, dWs), (rMs, rWs) = _sinks(ctx.params)
        table, n_tiles = ctx.plan.snet_table()
        T_v = float(ctx.plan.tokens)
        call("umpr_snet_bwd_tc", ptr(x), ptr(table), n_tiles, ptr(d_sa), ptr(Ms), ptr(Ws), N, L, ptr(dx), ptr(dMs), ptr(dWs), _n_ctas(dev),
             work=(6.0 * T_v * D * ATT, T_v * D * 8.0))
        d_word_soft = d_wsum.view(N, 1).expand(N, Wd).reshape(ws_shape) if want_ws else None
        if ctx.sink is not None and ctx.sink.armed and want_ws:
            ctx.sink.park(dx)                     # the co-attention backward (which needs d_word_soft, so it runs after us) adds it in
            dx = None
        return None, None, dx, d_word_soft, None, rMs, rWs


def s_net(gru_repr, word_soft, sent_length, Ms, Ws, plan=None, sink=None):
    """→ self_atte (B,S,128), sentiment (B,128).  ``plan``: the PackPlan of the ImprovedRnn call that produced ``gru_repr`` (its rows
    beyond each sentence's length are exactly zero) - enables the tensor-core kernels, which skip them."""
    if (TENSOR_CORE_SNET and plan is not None and plan.R == 128 and plan.L == int(sent_length) and plan.L <= 128
            and plan.N * plan.L == gru_repr.shape[0] * gru_repr.shape[1] and tuple(Ms.shape) == (ATT, D) and gru_repr.shape[2] == D):
        return _SNetTcFn.apply(plan, sink, gru_repr, word_soft, sent_length, Ms, Ws)
    return _SNetFn.apply(gru_repr, word_soft, sent_length, Ms, Ws)


# --------------------------------------------------------------------------------------------------------------------
# text matching   (model.py:166-168): tanh(linear_u([atte_u ‖ senti_u]) + linear_i([atte_i ‖ senti_i])), bias-free
# --------------------------------------------------------------------------------------------------------------------
class _TextMatchFn(Function):
    @staticmethod
    def forward(ctx, atte_u, senti_u, atte_i, senti_i, Wu, Wi):
        ctx.params = (Wu, Wi)
        ins = [_f32(t) for t in (atte_u, senti_u, atte_i, senti_i)]
        Wu, Wi = _f32(Wu), _f32(Wi)
        B = ins[0].shape[0]
        y = torch.empty(B, D, dtype=torch.float32, device=ins[0].device)
        if tuple(Wu.shape) != (D, 2 * D) or tuple(Wi.shape) != (D, 2 * D) or any(tuple(t.shape) != (B, D) for t in ins):
            raise RuntimeError("umpr_b200: text matching is built for gru_size=64 (two bias-free 256 -> 128 linears)")
        call("umpr_text_match_fwd", ptr(ins[0]), ptr(ins[1]), ptr(ins[2]), ptr(ins[3]), ptr(Wu), ptr(Wi), B, ptr(y),
             work=(2.0 * B * D * 4 * D, 0.0))
        ctx.save_for_backward(*ins, Wu, Wi, y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        a_u, s_u, a_i, s_i, Wu, Wi, y = ctx.saved_tensors
        B = y.shape[0]
        dev = y.device
        dpre = torch.empty_like(y)
        call("umpr_tanh_bwd", ptr(y), ptr(_f32(dy)), y.numel(), ptr(dpre))
        dins = torch.empty(4, B, D, dtype=torch.float32, device=dev)
        dW, rW = _sinks(ctx.params)
        call("umpr_text_match_bwd", ptr(dpre), ptr(Wu), ptr(Wi), B, ptr(dins[0]), ptr(dins[1]), ptr(dins[2]), ptr(dins[3]),
             work=(2.0 * B * D * 4 * D, 0.0))
        # dW[:, half] += dpre^T · in_j: reduction over the batch for all four halves in one launch
        call("umpr_text_match_wgrad", ptr(dpre), ptr(a_u), ptr(s_u), ptr(a_i), ptr(s_i), B, ptr(dW[0]), ptr(dW[1]),
             work=(2.0 * B * D * 4 * D, 0.0))
        return dins[0], dins[1], dins[2], dins[3], rW[0], rW[1]


def text_match(atte_u, senti_u, atte_i, senti_i, Wu, Wi):
    return _TextMatchFn.apply(atte_u, senti_u, atte_i, senti_i, Wu, Wi)


# --------------------------------------------------------------------------------------------------------------------
# CNet tail   (model.py:118-125)
# --------------------------------------------------------------------------------------------------------------------
class _CNetTailFn(Function):
    @staticmethod
    def forward(ctx, plan, gru_repr, sent_count, sent_length, conv_w, conv_b, lin_w, lin_b, threshold):
        ctx.params = (conv_w, conv_b, lin_w, lin_b)
        x = _f32(_chk(gru_repr, "gru_repr"))
        conv_w, conv_b, lin_w, lin_b = _f32(conv_w), _f32(conv_b), _f32(lin_w), _f32(lin_b)
        B = x.shape[0]
        S, L = int(sent_count), int(sent_length)
        N = B * S
        KC, cin, ks = conv_w.shape
        V = lin_w.shape[0]
        dev = x.device
        if cin != D:
            raise RuntimeError("umpr_b200: CNet conv is built for in_channels=128")
        cfeat = torch.empty(N, KC, dtype=torch.float32, device=dev)
        cidx = torch.empty(N, KC, dtype=torch.int32, device=dev)
        work = (2.0 * N * L * 3 * D * KC, N * L * D * 4.0)
        if TENSOR_CORE_CONV:
            cap = max(4096, (N * KC + 7) // 8)           # one 2-byte re-scoring record per (sentence, filter)
            scratch = torch.empty(_workspace_floats("cnet_conv_fwd_tc", cap), dtype=torch.float32, device=dev)
            table, n_tiles = None, 0
            ctx.keep = None
            if plan is not None and plan.N == N and plan.L == L and L + 2 <= 128:
                # gru_repr comes out of ImprovedRnn with this plan: rows beyond each sentence's length are exactly zero
                table, n_tiles = plan.cnet_table()
                ctx.keep = plan
                work = (2.0 * (plan.tokens + N) * 3 * D * KC, plan.tokens * D * 4.0)
            call("umpr_cnet_conv_fwd_tc", ptr(x), ptr(conv_w), ptr(conv_b), N, L, KC, ks, ptr(table), n_tiles, ptr(scratch), cap,
                 ptr(cfeat), ptr(cidx), _n_ctas(dev), work=work)
        else:
            wt = torch.empty(3 * D * 128, dtype=torch.float32, device=dev)
            call("umpr_cnet_prep", ptr(conv_w), KC, ks, ptr(wt))
            call("umpr_cnet_conv_fwd", ptr(x), ptr(wt), ptr(conv_b), N, L, KC, ptr(cfeat), ptr(cidx), _n_ctas(dev), work=work)
        if ROUTING_LOG is not None:
            ROUTING_LOG.append(("cnet", cidx.clone()))
        view_p = torch.empty(B, S, V, dtype=torch.float32, device=dev)
        final = torch.empty(B, V, dtype=torch.float32, device=dev)
        call("umpr_cnet_head_fwd", ptr(cfeat), ptr(lin_w), ptr(lin_b), float(threshold), B, S, V, KC, ptr(view_p), ptr(final))
        ctx.dims = (B, S, L, KC, V)
        ctx.save_for_backward(x, cfeat, cidx, view_p, conv_w, lin_w)
        return view_p, final

    @staticmethod
    @once_differentiable
    def backward(ctx, d_view_p, d_final):
        x, cfeat, cidx, view_p, conv_w, lin_w = ctx.saved_tensors
        B, S, L, KC, V = ctx.dims
        N = B * S
        dev = x.device
        dcfeat = torch.empty(N, KC, dtype=torch.float32, device=dev)
        (d_conv_w, d_conv_b, d_lin_w, d_lin_b), rets = _sinks(ctx.params)
        call("umpr_cnet_head_bwd", ptr(cfeat), ptr(cidx), ptr(view_p), ptr(lin_w), ptr(None if d_view_p is None else _f32(d_view_p)),
             ptr(None if d_final is None else _f32(d_final)), B, S, V, KC, ptr(dcfeat), ptr(d_lin_w), ptr(d_lin_b), ptr(d_conv_b))
        dx = _grad_like(x) if getattr(ctx, "keep", None) is not None else torch.empty_like(x)   # with a plan: rows beyond a sentence's length stay unwritten (never read, model.py:18)
        wt = torch.empty(_workspace_floats("cnet_conv_bwd_dx", KC), dtype=torch.float32, device=dev)
        cst = None
        plan = getattr(ctx, "keep", None)
        if plan is not None:
            table, n_tiles = plan.snet_table()
            cst = table.data_ptr() + 4 * (n_tiles + 1)
        rows = plan.tokens if plan is not None else N * L
        if plan is not None and TENSOR_CORE_CONV and TENSOR_CORE_CONV_DX:
            ctab, c_tiles = plan.cnet_table()
            wimg = torch.empty(_workspace_floats("cnet_conv_bwd_dx_tc", 0), dtype=torch.float32, device=dev)
            call("umpr_cnet_conv_bwd_dx_tc", ptr(dcfeat), ptr(cidx), ptr(conv_w), N, L, KC, ptr(ctab), c_tiles, ptr(wimg), ptr(dx), _n_ctas(dev),
                 work=(2.0 * (rows + 2 * N) * 3 * D * KC, rows * D * 4.0 + N * KC * 8.0))
        else:
            call("umpr_cnet_conv_bwd_dx", ptr(dcfeat), ptr(cidx), ptr(conv_w), N, L, KC, cst, ptr(wt), ptr(dx), _n_ctas(dev),
                 work=(2.0 * N * KC * 3 * D, rows * D * 4.0 + N * KC * 8.0))
        if plan is not None and TENSOR_CORE_CONV:
            ctab, c_tiles = plan.cnet_table()
            call("umpr_cnet_conv_bwd_dw_tc", ptr(x), ptr(dcfeat), ptr(cidx), N, L, KC, ptr(ctab), c_tiles, ptr(d_conv_w), _n_ctas(dev),
                 work=(2.0 * (rows + 2 * N) * 3 * D * KC, rows * D * 4.0 + N * KC * 8.0))
        else:
            call("umpr_cnet_conv_bwd_dw", ptr(x), ptr(dcfeat), ptr(cidx), N, L, KC, cst, ptr(d_conv_w), _n_ctas(dev),
                 work=(2.0 * N * KC * 3 * D, rows * D * 4.0 + N * KC * 8.0))
        return None, dx, None, None, rets[0], rets[1], rets[2], rets[3], None


def c_net_tail(gru_repr, sent_count, sent_length, conv_w, conv_b, lin_w, lin_b, threshold, plan=None):
    """→ view_p (B,S,V), final_repr (B,V).  ``plan``: the PackPlan of the ImprovedRnn call that produced ``gru_repr`` (rows beyond
    each sentence's length exactly zero) - the tensor-core convolution then lays out and multiplies the valid rows only."""
    return _CNetTailFn.apply(plan, gru_repr, sent_count, sent_length, conv_w, conv_b, lin_w, lin_b, threshold)


# --------------------------------------------------------------------------------------------------------------------
# ControlNet tail   (model.py:186-197; SSNet model.py:142-143)
# --------------------------------------------------------------------------------------------------------------------
class _ControlTailFn(Function):
    @staticmethod
    def forward(ctx, s, view_p, c_out, ss_w, ss_b, eps):
        ctx.params = (ss_w, ss_b)
        s, view_p, c_out, ss_w, ss_b = (_f32(t) for t in (s, view_p, c_out, ss_w, ss_b))
        B, Su, _ = s.shape
        V = view_p.shape[-1]
        dev = s.device
        senti = torch.empty(B, Su, dtype=torch.float32, device=dev)
        out = torch.empty(3, B, V, dtype=torch.float32, device=dev)       # score, prefer_pos, prefer_neg
        call("umpr_control_tail_fwd", ptr(s), ptr(view_p), ptr(c_out), ptr(ss_w), ptr(ss_b), float(eps), B, Su, V, ptr(senti), ptr(out[0]),
             ptr(out[1]), ptr(out[2]))
        ctx.eps = float(eps)
        ctx.save_for_backward(s, view_p, c_out, ss_w, senti, out)
        return out[1], out[2]

    @staticmethod
    @once_differentiable
    def backward(ctx, d_pp, d_pn):
        s, view_p, c_out, ss_w, senti, out = ctx.saved_tensors
        B, Su, _ = s.shape
        V = view_p.shape[-1]
        dev = s.device
        z = lambda t: torch.zeros(B, V, dtype=torch.float32, device=dev) if t is None else _f32(t)
        d_s = torch.empty_like(s)
        d_vp = torch.empty_like(view_p)
        d_co = torch.empty_like(c_out)
        (dw_, db_), (rw_, rb_) = _sinks(ctx.params)
        call("umpr_control_tail_bwd", ptr(s), ptr(view_p), ptr(c_out), ptr(ss_w), ptr(senti), ptr(out[0]), ptr(z(d_pp)), ptr(z(d_pn)),
             ctx.eps, B, Su, V, ptr(d_s), ptr(d_vp), ptr(d_co), ptr(dw_), ptr(db_))
        return d_s, d_vp, d_co, rw_, rb_, None


class _SSNetFn(Function):
    """Standalone SSNet (model.py:142-143): sigmoid(Linear(128 -> 1)).  Inside UMPR it is fused into the ControlNet tail."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.params = (w, b)
        ctx.shape = tuple(x.shape)
        x, w, b = _f32(_chk(x, "sentiment_emb")), _f32(w), _f32(b)
        if x.shape[-1] != D or w.numel() != D:
            raise RuntimeError("umpr_b200: SSNet is built for input_size=128 (2 * gru_size)")
        rows = x.numel() // D
        y = torch.empty(rows, dtype=torch.float32, device=x.device)
        call("umpr_ssnet_fwd", ptr(x), ptr(w), ptr(b), rows, ptr(y))
        ctx.save_for_backward(x, w, y)
        return y.view(*ctx.shape[:-1], 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        rows = y.numel()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        (dw, db), (rw, rb) = _sinks(ctx.params)
        call("umpr_ssnet_bwd", ptr(x), ptr(w), ptr(y), ptr(_f32(dy).reshape(-1)), rows, ptr(dx), ptr(dw), ptr(db))
        return (dx.view(ctx.shape) if dx is not None else None), rw, rb


def ss_net(x, w, b):
    """→ sigmoid(x @ w^T + b), shape (..., 1)."""
    return _SSNetFn.apply(x, w, b)


def control_tail(s, view_p, c_out, ss_w, ss_b, eps):
    """→ prefer_pos, prefer_neg (B,V)."""
    return _ControlTailFn.apply(s, view_p, c_out, ss_w, ss_b, eps)


# --------------------------------------------------------------------------------------------------------------------
# VisualNet tail   (model.py:219-228)
# --------------------------------------------------------------------------------------------------------------------
class _VisualFn(Function):
    @staticmethod
    def forward(ctx, feat, c_u, c_i, pos_e, neg_e, w, b):
        ctx.params = (w, b)
        feat, c_u, c_i, pos_e, neg_e, w, b = (_f32(t) for t in (feat, c_u, c_i, pos_e, neg_e, w, b))
        _chk(feat, "photo features")
        B, V, Pc, Fd = feat.shape
        dev = feat.device
        emb = torch.empty(2, V, dtype=torch.float32, device=dev)
        out = torch.empty(5, B, V, dtype=torch.float32, device=dev)       # img_emb, pos_match, neg_match, final_pos, final_neg
        call("umpr_visual_fwd", ptr(feat), ptr(pos_e), ptr(neg_e), ptr(w), ptr(b), ptr(c_u), ptr(c_i), B, V, Pc, Fd, ptr(emb), ptr(out[0]),
             ptr(out[1]), ptr(out[2]), ptr(out[3]), ptr(out[4]))
        ctx.save_for_backward(feat, c_u, c_i, pos_e, neg_e, w, emb, out)
        return out[1], out[2], out[3], out[4]

    @staticmethod
    @once_differentiable
    def backward(ctx, d_pm, d_nm, d_fp, d_fn):
        feat, c_u, c_i, pos_e, neg_e, w, emb, out = ctx.saved_tensors
        B, V, Pc, Fd = feat.shape
        dev = feat.device
        c = lambda t: None if t is None else _f32(t)
        scratch = torch.empty(3 * B * V, dtype=torch.float32, device=dev)
        d_c = torch.empty(2, B, V, dtype=torch.float32, device=dev)
        d_e = torch.empty(2, V, Fd, dtype=torch.float32, device=dev)
        (dw_, db_), (rw_, rb_) = _sinks(ctx.params)
        call("umpr_visual_bwd", ptr(feat), ptr(pos_e), ptr(neg_e), ptr(w), ptr(emb), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(c_u),
             ptr(c_i), ptr(c(d_pm)), ptr(c(d_nm)), ptr(c(d_fp)), ptr(c(d_fn)), B, V, Pc, Fd, ptr(scratch), ptr(d_c[0]), ptr(d_c[1]),
             ptr(d_e[0]), ptr(d_e[1]), ptr(dw_), ptr(db_))
        return None, d_c[0], d_c[1], d_e[0], d_e[1], rw_, rb_


def visual_tail(feat, c_u, c_i, pos_e, neg_e, w, b):
    """→ pos_match, neg_match, final_pos, final_neg (B,V)."""
    return _VisualFn.apply(feat, c_u, c_i, pos_e, neg_e, w, b)


# --------------------------------------------------------------------------------------------------------------------
# fusion + losses   (model.py:268-277)
# --------------------------------------------------------------------------------------------------------------------
class _FusionFn(Function):
    @staticmethod
    def forward(ctx, repr_, fpos, fneg, w, b):
        ctx.params = (w, b)
        repr_, w, b = _f32(_chk(repr_, "review_net_repr")), _f32(w), _f32(b)
        fpos = None if fpos is None else _f32(fpos)
        fneg = None if fneg is None else _f32(fneg)
        B = repr_.shape[0]
        V = 0 if fpos is None else fpos.shape[1]
        pred = torch.empty(B, dtype=torch.float32, device=repr_.device)
        call("umpr_fusion_fwd", ptr(repr_), ptr(fpos), ptr(fneg), ptr(w), ptr(b), B, V, ptr(pred))
        ctx.V = V
        ctx.save_for_backward(repr_, fpos, fneg, w, pred)
        return pred

    @staticmethod
    @once_differentiable
    def backward(ctx, d_pred):
        repr_, fpos, fneg, w, pred = ctx.saved_tensors
        B, V = repr_.shape[0], ctx.V
        dev = repr_.device
        d_repr = torch.empty_like(repr_)
        d_f = torch.empty(2, B, V, dtype=torch.float32, device=dev) if V else None
        (dw_, db_), (rw_, rb_) = _sinks(ctx.params)
        call("umpr_fusion_bwd", ptr(repr_), ptr(fpos), ptr(fneg), ptr(w), ptr(pred), ptr(_f32(d_pred)), B, V, ptr(d_repr),
             ptr(d_f[0]) if V else None, ptr(d_f[1]) if V else None, ptr(dw_), ptr(db_))
        return d_repr, (d_f[0] if V else None), (d_f[1] if V else None), rw_, rb_


def fusion(repr_, fpos, fneg, w, b):
    """relu(Linear(cat[repr, final_pos, final_neg])).squeeze(-1) → (B,)"""
    return _FusionFn.apply(repr_, fpos, fneg, w, b)


class _LossFn(Function):
    @staticmethod
    def forward(ctx, pred, labels, pp, pn, pm, nm, rate):
        pred, labels = _f32(_chk(pred, "prediction")), _f32(labels)
        full = pp is not None
        if full:
            pp, pn, pm, nm = (_f32(t) for t in (pp, pn, pm, nm))
        B = pred.shape[0]
        V = pp.shape[1] if full else 0
        loss = torch.zeros(1, dtype=torch.float32, device=pred.device)
        call("umpr_loss_fwd", ptr(pred), ptr(labels), ptr(pp), ptr(pn), ptr(pm), ptr(nm), B, V, float(rate), ptr(loss))
        ctx.rate, ctx.V = float(rate), V
        ctx.save_for_backward(pred, labels, pp, pn, pm, nm)
        return loss.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        pred, labels, pp, pn, pm, nm = ctx.saved_tensors
        B, V = pred.shape[0], ctx.V
        dev = pred.device
        d_pred = torch.empty_like(pred)
        g = torch.empty(4, B, V, dtype=torch.float32, device=dev) if V else None
        gp = (lambda i: ptr(g[i])) if V else (lambda i: None)
        call("umpr_loss_bwd", ptr(pred), ptr(labels), ptr(pp), ptr(pn), ptr(pm), ptr(nm), ptr(_f32(d_loss).reshape(1)), B, V, ctx.rate,
             ptr(d_pred), gp(0), gp(1), gp(2), gp(3))
        if V:
            return d_pred, None, g[0], g[1], g[2], g[3], None
        return d_pred, None, None, None, None, None, None


def umpr_loss(pred, labels, pp=None, pn=None, pm=None, nm=None, rate=0.0):
    """mse_loss(pred, labels, 'mean') [+ rate * mean(pp^T @ pm + pn^T @ nm)]  (model.py:269,275-277)"""
    return _LossFn.apply(pred, labels, pp, pn, pm, nm, rate)


# --------------------------------------------------------------------------------------------------------------------
# R-Net pre-training head   (pretrain/pretrain_rnet.py:148-168)
# --------------------------------------------------------------------------------------------------------------------
class _BceHeadFn(Function):
    @staticmethod
    def forward(ctx, att_u, att_i, w, b, target):
        ctx.params = (w, b)
        att_u, att_i, w, b, target = (_f32(_chk(t, "bce head input")) for t in (att_u, att_i, w, b, target))
        B = att_u.shape[0]
        if att_u.shape[1] != D or w.numel() != 2 * D:
            raise RuntimeError("umpr_b200: the pre-training head is built for gru_size=64 (256 -> 1)")
        result = torch.empty(B, dtype=torch.float32, device=att_u.device)
        loss = torch.zeros((), dtype=torch.float32, device=att_u.device)
        call("umpr_bce_head_fwd", ptr(att_u), ptr(att_i), ptr(w), ptr(b), ptr(target), B, ptr(result), ptr(loss))
        ctx.save_for_backward(att_u, att_i, w, result, target)
        return result, loss

    @staticmethod
    @once_differentiable
    def backward(ctx, d_result, d_loss):
        att_u, att_i, w, result, target = ctx.saved_tensors
        B = att_u.shape[0]
        d_u, d_i = torch.empty_like(att_u), torch.empty_like(att_i)
        (dw, db), (rw, rb) = _sinks(ctx.params)
        call("umpr_bce_head_bwd", ptr(att_u), ptr(att_i), ptr(w), ptr(result), ptr(target), ptr(None if d_loss is None else _f32(d_loss)),
             ptr(None if d_result is None else _f32(d_result)), B, ptr(d_u), ptr(d_i), ptr(dw), ptr(db))
        return d_u, d_i, rw, rb, None


def bce_head(att_u, att_i, w, b, target):
    """→ result (B,), loss scalar."""
    return _BceHeadFn.apply(att_u, att_i, w, b, target)
