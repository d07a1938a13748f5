"""Development aid: CUDA-event time of single entry points on a music_full-shaped review side (batch 1024: 20480 sentences of <= 20
tokens, valid-row tables), L2 flushed between calls.  UMPR_DBG=<bits> switches roles of the tile kernels off (csrc: dbg_flags).
    python profiles/dbg_time.py ws snet_fwd snet_bwd dx dw tn"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from umpr_b200 import _lib, synthetic as syn  # noqa: E402
from umpr_b200._lib import call, ptr  # noqa: E402
from umpr_b200.functional import _workspace_floats  # noqa: E402
from umpr_b200.plan import PackPlan  # noqa: E402

dev = torch.device("cuda", 0)
_lib.load()
B, S, L, D, KC = 1024, 20, 20, 128, 120
N = B * S
rs = np.random.RandomState(0)
_, lens = syn._side(rs, B, S, L, 1000, 5, False, False)
plan = PackPlan(lens.reshape(-1), L, dev, tile_rows=128)
lens_row = plan.row_lengths().to(dev)
mask = (torch.arange(L, device=dev)[None, :] < lens_row[:, None])
x = (torch.tanh(torch.randn(N, L, D, device=dev)) * mask[:, :, None]).contiguous()
stab, s_tiles = plan.snet_table()
ctab, c_tiles = plan.cnet_table()
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
n_ctas = _lib.sm_count(dev)
print(f"tokens {int(lens.sum())}, snet tiles {s_tiles}, cnet tiles {c_tiles}, UMPR_DBG={os.environ.get('UMPR_DBG', '0')}")


def timed(name, fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"  {name:10s} {sorted(ts)[len(ts) // 2] * 1e3:8.1f} us")


what = sys.argv[1:] or ["ws", "snet_fwd", "snet_bwd", "dx", "dw", "tn"]
M = torch.randn(D, D, device=dev) * 0.1
Ms, Ws = torch.randn(64, D, device=dev) * 0.1, torch.randn(64, device=dev) * 0.1
out = torch.zeros(N * L, D, device=dev)
if "ws" in what:
    timed("ws", lambda: call("umpr_tc_gemm_ws", ptr(x), D, ptr(M), D, ptr(out), D, N * L, D, D, 0, None, 0, 1, ptr(stab), s_tiles, L, n_ctas))
    timed("ws(acc)", lambda: call("umpr_tc_gemm_ws", ptr(x), D, ptr(M), D, ptr(out), D, N * L, D, D, 1, None, 0, 0, ptr(stab), s_tiles, L, n_ctas))
sa = torch.empty(N, D, device=dev)
if "snet_fwd" in what:
    timed("snet_fwd", lambda: call("umpr_snet_fwd_tc", ptr(x), ptr(stab), s_tiles, ptr(Ms), ptr(Ws), N, L, ptr(sa), n_ctas))
if "snet_bwd" in what:
    d_sa = torch.randn(N, D, device=dev)
    dMs, dWs = torch.zeros(64, D, device=dev), torch.zeros(64, device=dev)
    timed("snet_bwd", lambda: call("umpr_snet_bwd_tc", ptr(x), ptr(stab), s_tiles, ptr(d_sa), ptr(Ms), ptr(Ws), N, L, ptr(out), ptr(dMs), ptr(dWs), n_ctas))
if "dx" in what or "dw" in what:
    conv_w = torch.randn(KC, D, 3, device=dev) * 0.05
    dcfeat = torch.randn(N, KC, device=dev)
    cidx = (torch.rand(N, KC, device=dev) * (lens_row[:, None].float() + 1)).to(torch.int32).clamp(max=L - 1)
    cidx = torch.where(torch.rand(N, KC, device=dev) < 0.3, torch.full_like(cidx, -1), cidx)
    if "dx" in what:
        wimg = torch.empty(_workspace_floats("cnet_conv_bwd_dx_tc", 0), dtype=torch.float32, device=dev)
        timed("conv_dx", lambda: call("umpr_cnet_conv_bwd_dx_tc", ptr(dcfeat), ptr(cidx), ptr(conv_w), N, L, KC, ptr(ctab), c_tiles, ptr(wimg), ptr(out), n_ctas))
    if "dw" in what:
        dW = torch.zeros(KC, D, 3, device=dev)
        timed("conv_dw", lambda: call("umpr_cnet_conv_bwd_dw_tc", ptr(x), ptr(dcfeat), ptr(cidx), N, L, KC, ptr(ctab), c_tiles, ptr(dW), n_ctas))
if "tn" in what:
    dM = torch.zeros(D, D, device=dev)
    y = torch.randn(N * L, D, device=dev)
    timed("gemm_tn", lambda: call("umpr_tc_gemm_tn", ptr(x), D, ptr(y), D, ptr(dM), D, D, D, N * L, n_ctas))
if "coattn" in what:
    from umpr_b200 import functional as F
    _, lens_i = syn._side(rs, B, S, L, 1000, 5, False, False)
    plan_i = PackPlan(lens_i.reshape(-1), L, dev, tile_rows=128)
    mask_i = (torch.arange(L, device=dev)[None, :] < plan_i.row_lengths().to(dev)[:, None])
    gi = (torch.tanh(torch.randn(N, L, D, device=dev)) * mask_i[:, :, None]).view(B, S * L, D).contiguous()
    gu = x.view(B, S * L, D)
    with torch.no_grad():
        timed("coattn_fwd", lambda: F.co_attention(gu, gi, M, plans=(plan, plan_i)))
