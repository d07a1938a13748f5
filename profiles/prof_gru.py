"""Fused tensor-core GRU launches at bench shapes (B=1024, S=20, L=20; user+item = one launch) for CUDA-event timing and ncu.

    python profiles/prof_gru.py [fwd|fwd_infer|bwd ...]
    ncu --set full --clock-control none --import-source on -k regex:gru_.*_tc_kernel -c 2 -o gpurun_out/prof_gru python profiles/prof_gru.py fwd
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import functional as F  # noqa: E402
from umpr_b200 import synthetic as syn  # noqa: E402
from umpr_b200.model import PackedReviews  # noqa: E402

ops = sys.argv[1:] or ["fwd", "fwd_infer", "bwd"]
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 1024))
torch.manual_seed(0)
table = syn.make_table(400003).to(dev)
batch = syn.make_batch("music_full", B, seed=0)
gru = torch.nn.GRU(50, 64, batch_first=True, bidirectional=True).to(dev)
w = [gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse,
     gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse]
pu = PackedReviews(batch[3], ids=batch[0].to(dev), table=table)
pi = PackedReviews(batch[4], ids=batch[1].to(dev), table=table)
tokens = pu.plan.tokens + pi.plan.tokens
slots = pu.plan.slots + pi.plan.slots


def timed(fn, reps=5):
    """CUDA events around each C-ABI call only (host-side planning excluded) -> best and mean of the summed kernel time."""
    from umpr_b200 import _lib
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        _lib.start_timing()
        fn()
        tab = _lib.stop_timing()
        ms.append(sum(v["ms"] for v in tab.values()))
    print("   ", {k: round(v["ms"], 3) for k, v in tab.items()})
    return min(ms), sum(ms) / len(ms)


def fwd(train):
    ws = w if train else [t.detach() for t in w]
    with torch.set_grad_enabled(train):
        return F.gru_forward_multi([pu.plan, pi.plan], [pu.xp, pi.xp], [pu.xq, pi.xq], pu.E, ws, False)


print(f"B={B}: {tokens} valid tokens, {slots} computed token slots, tiles {pu.plan.n_tiles}+{pi.plan.n_tiles}")
if "fwd" in ops:
    best, mean = timed(lambda: fwd(True))
    print(f"fwd (train, user+item one launch incl. python): best {best:.3f} ms mean {mean:.3f} ms -> {2.0 * tokens * 114 * 384 / best / 1e9:.1f} algorithmic TFLOP/s")
if "fwd_infer" in ops:
    best, mean = timed(lambda: fwd(False))
    print(f"fwd (inference): best {best:.3f} ms mean {mean:.3f} ms -> {2.0 * tokens * 114 * 384 / best / 1e9:.1f} algorithmic TFLOP/s")
if "bwd" in ops:
    (gu, _), (gi, _) = fwd(True)
    gy = torch.randn_like(gu)

    def bwd():
        for t in w:
            t.grad = None
        torch.autograd.backward([gu, gi], [gy, gy], retain_graph=True)
    best, mean = timed(bwd)
    print(f"bwd (recurrence + wgrad, both sides): best {best:.3f} ms mean {mean:.3f} ms")
torch.cuda.synchronize()
print("done", ops)
