"""Isolated launches of the big kernels at bench shapes (B=1024, S=20, L=20) for ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ops.csv python profiles/prof_ops.py [op ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import functional as F  # noqa: E402
from umpr_b200 import synthetic as syn  # noqa: E402
from umpr_b200.model import PackedReviews  # noqa: E402

ops = sys.argv[1:] or ["coattn", "gru", "snet", "cnet"]
dev = torch.device("cuda", 0)
B, S, L = 1024, 20, 20
torch.manual_seed(0)
table = syn.make_table(400003).to(dev)
batch = syn.make_batch("music_full", B, seed=0)
gru = torch.nn.GRU(50, 64, batch_first=True, bidirectional=True).to(dev)
w = [gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse,
     gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse]
pu = PackedReviews(batch[3], ids=batch[0].to(dev), table=table)
pi = PackedReviews(batch[4], ids=batch[1].to(dev), table=table)
for rep in range(2):
    gu, _ = F.gru_forward(pu.plan, pu.xp, pu.E, w, False)
    gi, _ = F.gru_forward(pi.plan, pi.xp, pi.E, w, False)
    gu = gu.view(B, S * L, 128)
    gi = gi.view(B, S * L, 128)
    if "gru" in ops:
        (gu.sum() + gi.sum()).backward()
    if "coattn" in ops:
        M = (torch.randn(128, 128, device=dev)).requires_grad_(True)
        gud, gid = gu.detach().requires_grad_(True), gi.detach().requires_grad_(True)
        su, si, au, ai = F.co_attention(gud, gid, M)
        (su.sum() + si.sum() + au.sum() + ai.sum()).backward()
    if "snet" in ops:
        Ms = torch.randn(64, 128, device=dev, requires_grad=True)
        Ws = torch.randn(1, 64, device=dev, requires_grad=True)
        gud = gu.detach().requires_grad_(True)
        sa, se = F.s_net(gud, torch.rand(B, S * L, device=dev), L, Ms, Ws)
        (sa.sum() + se.sum()).backward()
    if "cnet" in ops:
        cw = (torch.randn(120, 128, 3, device=dev) * 0.05).requires_grad_(True)
        cb = torch.zeros(120, device=dev, requires_grad=True)
        lw = torch.randn(1, 120, device=dev, requires_grad=True)
        lb = torch.zeros(1, device=dev, requires_grad=True)
        gud = gu.detach().requires_grad_(True)
        vp, fr = F.c_net_tail(gud, S, L, cw, cb, lw, lb, 0.35)
        (vp.sum() + fr.sum()).backward()
torch.cuda.synchronize()
print("done", ops)
