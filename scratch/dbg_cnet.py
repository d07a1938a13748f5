import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpr_b200 import synthetic as syn, functional as F
DEV = "cuda:0"
WL, BB, seed = sys.argv[1].split(":")
table = syn.make_table(3000, seed=2)
bt = syn.make_batch(WL, int(BB), vocab=3000, seed=int(seed))
model = syn.build_model(WL, table, seed=1, device=DEV)
with torch.no_grad():
    model.review_net.r_net.M.mul_(0.05)
model.train()
caps = []
orig = F.c_net_tail
def hook(x, S, L, cw, cb, lw, lb, thr):
    r = orig(x, S, L, cw, cb, lw, lb, thr)
    caps.append((x.detach(), S, L, cw.detach(), cb.detach(), lw.detach(), lb.detach(), thr, r[0].detach()))
    return r
F.c_net_tail = hook
pred, loss = model(*bt)
for x, S, L, cw, cb, lw, lb, thr, vp in caps:
    B = x.shape[0]
    xs = x.view(B * S, L, 128).transpose(-1, -2)
    c64 = torch.relu(torch.nn.functional.conv1d(xs.double(), cw.double(), cb.double(), padding=1))
    c32 = torch.relu(torch.nn.functional.conv1d(xs, cw, cb, padding=1))
    top2 = c64.topk(2, dim=-1)
    gap = (top2.values[..., 0] - top2.values[..., 1])
    pos = top2.values[..., 0] > 0
    rel = torch.where(pos, gap / top2.values[..., 0].clamp_min(1e-30), torch.ones_like(gap))
    a64, a32 = c64.argmax(-1), c32.argmax(-1)
    print("side N=%d: min rel gap %.3e, #gaps<1e-6: %d, #exact ties (positive max): %d, fp32/fp64 torch argmax differ: %d" % (
        B * S, float(rel.min()), int((rel < 1e-6).sum()), int(((gap == 0) & pos).sum()), int(((a64 != a32) & pos).sum())))
    feat = c64.max(-1)[0].view(B, S, -1)
    v64 = torch.sigmoid(feat @ lw.double().t() + lb.double())
    print("   view_p min |v-thr| %.3e ; ours vs fp64 view_p max diff %.3e" % (float((v64 - thr).abs().min()),
          float((torch.where(v64 < thr, torch.zeros_like(v64), v64) - vp.double()).abs().max())))
F.c_net_tail = orig
for x0, S, L, cw, cb, lw, lb, thr, _ in caps:
    B, V = x0.shape[0], lw.shape[0]
    torch.manual_seed(0)
    gv, gf = torch.randn(B, S, V, device=DEV), torch.randn(B, V, device=DEV)
    x = x0.clone().requires_grad_(True)
    p = [t.clone().requires_grad_(True) for t in (cw, cb, lw, lb)]
    view_p, final = F.c_net_tail(x, S, L, *p, thr)
    ((view_p * gv).sum() + (final * gf).sum()).backward()
    got = [view_p.detach(), final.detach(), x.grad] + [t.grad for t in p]
    xd = x0.double().cpu().requires_grad_(True)
    pd = [t.double().cpu().requires_grad_(True) for t in (cw, cb, lw, lb)]
    conv = torch.relu(torch.nn.functional.conv1d(xd.view(B * S, L, 128).transpose(-1, -2), pd[0], pd[1], padding=1))
    feat = conv.max(dim=-1)[0].reshape(B, S, -1)
    vp = torch.sigmoid(feat @ pd[2].t() + pd[3])
    vp = torch.where(vp < thr, torch.zeros_like(vp), vp)
    fin = (vp ** 2).sum(-2)
    ((vp * gv.double().cpu()).sum() + (fin * gf.double().cpu()).sum()).backward()
    ref = [vp.detach(), fin.detach(), xd.grad] + [t.grad for t in pd]
    valid = (x0.abs().sum(-1, keepdim=True) > 0).cpu()
    out = []
    for a, b, nm in zip(got, ref, ["view_p", "final", "dx(valid rows)", "d conv_w", "d conv_b", "d lin_w", "d lin_b"]):
        a = a.double().cpu()
        if nm.startswith("dx"):
            a, b = a * valid, b * valid
        out.append("%s %.2e" % (nm, float((a - b).abs().max() / b.abs().max())))
    print("isolated tail, N=%d:" % (B * S), " | ".join(out), flush=True)
