"""Debug aid: the convolution forward inside the real model (music_full), checked per C-Net call against torch conv1d."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import _lib, functional as F, synthetic as syn  # noqa: E402
from umpr_b200._lib import call, ptr  # noqa: E402
from umpr_b200.functional import _workspace_floats  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
table = syn.make_table(50000, seed=2)
batch = syn.make_batch("music_full", B, vocab=50000, seed=5)
model = syn.build_model("music_full", table, seed=1, device=dev)
calls = []
orig = F.c_net_tail


def spy(gru_repr, S, L, cw, cb, lw, lb, thr, plan=None):
    calls.append((gru_repr.detach().clone(), S, L, plan))
    return orig(gru_repr, S, L, cw, cb, lw, lb, thr, plan=plan)


F.c_net_tail = spy
import umpr_b200.model as M  # noqa: E402
M.F.c_net_tail = spy
model.train()
model(*batch)
cw, cb = model.control_net.c_net.cnn[0].weight.detach(), model.control_net.c_net.cnn[0].bias.detach()
KC = cw.shape[0]
print("calls", len(calls), "bias range", float(cb.min()), float(cb.max()))
for x, S, L, plan in calls:
    N = x.shape[0] * S
    xs = x.reshape(N, L, 128).contiguous()
    cap = max(4096, (N * KC + 7) // 8)
    scratch = torch.empty(_workspace_floats("cnet_conv_fwd_tc", cap), dtype=torch.float32, device=dev)
    cfeat = torch.full((N, KC), -7.0, dtype=torch.float32, device=dev)
    cidx = torch.full((N, KC), -9, dtype=torch.int32, device=dev)
    tbl, nt = plan.cnet_table() if plan is not None else (None, 0)
    call("umpr_cnet_conv_fwd_tc", ptr(xs), ptr(cw), ptr(cb), N, L, KC, 3, ptr(tbl), nt, ptr(scratch), cap, ptr(cfeat), ptr(cidx), 148)
    torch.cuda.synchronize()
    counter = (scratch.view(torch.int16)[(196608 + 1024) // 2:(196608 + 1024) // 2 + N * KC] >= 0).sum()
    y = torch.nn.functional.conv1d(xs.transpose(1, 2).double(), cw.double(), cb.double(), padding=1).float()
    top = y.max(dim=2).values
    want = torch.relu(top)
    err = (cfeat - want).abs()
    at = torch.where(cidx >= 0, torch.gather(y, 2, cidx.long().clamp(min=0)[:, :, None]).squeeze(2), torch.zeros_like(top))
    marg = (want - at).abs()
    lens = plan.row_lengths().cpu() if plan is not None else None
    zero_beyond = True
    if lens is not None:
        mask = (torch.arange(L)[None, :] >= lens[:, None]).to(dev)
        zero_beyond = bool((xs.abs().amax(dim=2)[mask] == 0).all())
    print(f"N={N} S={S} L={L} tiles={nt} worklist={int(counter)} cap={cap}: value err {float(err.max()):.3e}, pick margin {float(marg.max()):.3e} (rel {float(marg.max() / want.max()):.3e}), rows beyond len zero: {zero_beyond}")
    if float(marg.max()) > 1e-4:
        flat = int(marg.argmax())
        n, f = divmod(flat, KC)
        print(f"   worst: sentence {n} len {int(lens[n]) if lens is not None else L} filter {f}: picked {int(cidx[n, f])} cfeat {float(cfeat[n, f]):.6f} | pre at pick {float(at[n, f]):.6f}, "
              f"max {float(top[n, f]):.6f} at {int(y[n, f].argmax())}, bias {float(cb[f]):.6f}")
        print("   pre row:", [round(v, 5) for v in y[n, f].tolist()])
        nb = (marg > 1e-4).sum()
        print("   bad picks:", int(nb), "bad lens:", sorted(set(lens[(marg > 1e-4).any(dim=1).cpu()].tolist()))[:30] if lens is not None else None)
