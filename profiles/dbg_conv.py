"""Debug aid: the tcgen05 convolution forward (umpr_cnet_conv_fwd_tc) against torch conv1d on the GPU, mismatches located by
tile / pair / position.   python profiles/dbg_conv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import _lib  # noqa: E402
from umpr_b200._lib import call, ptr  # noqa: E402
from umpr_b200.functional import _workspace_floats  # noqa: E402
from umpr_b200.plan import PackPlan  # noqa: E402

dev = torch.device("cuda", 0)
_lib.load()


def ref(x, w, b, lens, L):
    N = x.shape[0]
    y = torch.nn.functional.conv1d(x.transpose(1, 2).double(), w.double(), b.double(), padding=1)      # (N, KC, L)
    return y.float()


def run(N, L, KC, n_ctas, use_plan, seed=0, short=False):
    torch.manual_seed(seed)
    lens = torch.randint(1, (min(3, L) if short else L) + 1, (N,))
    x = torch.tanh(torch.randn(N, L, 128, device=dev))
    mask = torch.arange(L)[None, :] < lens[:, None]
    x = x * mask.to(dev)[:, :, None]
    w = torch.randn(KC, 128, 3, device=dev) * 0.05
    b = torch.randn(KC, device=dev) * 0.3
    cap = max(4096, (N * KC + 7) // 8)
    scratch = torch.empty(_workspace_floats("cnet_conv_fwd_tc", cap), dtype=torch.float32, device=dev)
    cfeat = torch.full((N, KC), -7.0, dtype=torch.float32, device=dev)
    cidx = torch.full((N, KC), -9, dtype=torch.int32, device=dev)
    table, n_tiles = None, 0
    if use_plan:
        plan = PackPlan(lens, L, dev, tile_rows=128)
        table, n_tiles = plan.cnet_table()
        lens = plan.row_lengths().cpu()              # the reference's double un-sort: output row n carries another sentence's length
        mask = torch.arange(L)[None, :] < lens[:, None]
        x = x * mask.to(dev)[:, :, None]
    call("umpr_cnet_conv_fwd_tc", ptr(x), ptr(w), ptr(b), N, L, KC, 3, ptr(table), n_tiles, ptr(scratch), cap, ptr(cfeat), ptr(cidx), n_ctas)
    torch.cuda.synchronize()
    y = ref(x, w, b, lens, L)
    want = torch.relu(y.max(dim=2).values)
    err = (cfeat - want).abs()
    bad = err > 2e-4
    print(f"N={N} L={L} KC={KC} ctas={n_ctas} plan={use_plan} tiles={n_tiles}: max err {float(err.max()):.3e}, bad {int(bad.sum())} of {bad.numel()}")
    # the routing: value at the chosen position against the maximum
    pick = cidx.long().clamp(min=0)
    at = torch.relu(torch.gather(y, 2, pick[:, :, None]).squeeze(2)) * (cidx >= 0)
    marg = want - at
    print(f"   worst pick margin {float(marg.max()):.3e}; picks outside [0,L): {int(((cidx >= L) | (cidx < -1)).sum())}")
    if float(marg.max()) > 1e-4:
        n, f = divmod(int(marg.argmax()), KC)
        print(f"   worst pick: sentence {n} (len {int(lens[n])}) filter {f}: picked {int(cidx[n, f])} value {float(at[n, f]):.5f}, max {float(want[n, f]):.5f} at {int(y[n, f].argmax())}")
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten().cpu()
        print("   bad sentences:", rows[:40].tolist(), "...", int(rows.numel()))
        if use_plan:
            tso = table[:n_tiles + 1].cpu()
            tiles = torch.bucketize(rows, tso, right=True) - 1
            print("   their tiles:", sorted(set(tiles.tolist()))[:40])
            g = max(1, min(n_tiles, n_ctas))
            print("   local index of those tiles in their CTA:", sorted(set((tiles // g).tolist()))[:40])
        fb = bad.any(dim=0).nonzero().flatten().cpu()
        print("   bad filters:", fb[:20].tolist(), "...", int(fb.numel()))
        r0 = int(rows[0])
        print("   first bad row: len", int(lens[r0]), "got", cfeat[r0, :6].tolist(), "want", want[r0, :6].tolist())


for args in [] if "--time" in sys.argv else [(40, 20, 120, 148, True), (600, 20, 120, 4, True), (3000, 20, 120, 7, True),
             (20480, 20, 120, 148, True), (20480, 20, 120, 148, True, 1), (20480, 20, 120, 148, True, 2), (300, 100, 120, 3, True), (2000, 5, 100, 2, True)]:
    run(*args)
    run(*args, short=True)


def timeit(N=20480, L=20, KC=120):
    """kernel time of the convolution forward on a music_full-shaped side (lengths 6..20, sorted pools do not matter here)"""
    torch.manual_seed(0)
    lens = torch.randint(6, L + 1, (N,))
    plan = PackPlan(lens, L, dev, tile_rows=128)
    table, n_tiles = plan.cnet_table()
    x = torch.tanh(torch.randn(N, L, 128, device=dev))
    w = torch.randn(KC, 128, 3, device=dev) * 0.05
    b = torch.randn(KC, device=dev) * 0.05
    cap = max(4096, (N * KC + 7) // 8)
    scratch = torch.empty(_workspace_floats("cnet_conv_fwd_tc", cap), dtype=torch.float32, device=dev)
    cfeat = torch.empty(N, KC, device=dev)
    cidx = torch.empty(N, KC, dtype=torch.int32, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("umpr_cnet_conv_fwd_tc", ptr(x), ptr(w), ptr(b), N, L, KC, 3, ptr(table), n_tiles, ptr(scratch), cap, ptr(cfeat), ptr(cidx), 148)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"UMPR_CONV_DBG={os.environ.get('UMPR_CONV_DBG', '0')}: N={N} tiles={n_tiles}: {sorted(ts)[len(ts) // 2] * 1e3:.1f} us (prep + conv + fix)")


if "--time" in sys.argv:
    timeit()
