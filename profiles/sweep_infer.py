"""BASELINE.json configs[4]: test_only inference sweep (evaluate.py:6-14 — eval(), no_grad, forward only) over batch size x
padded sentence length with the length-skew stress of SURVEY.md §8d (90 % of the sentences 1-4 tokens, 10 % full length).

    python profiles/sweep_infer.py [--workload music_full] > gpurun_out/sweep.json

Runs through ``umpr_b200.evaluate_mse`` (the package's evaluate.py:6-14): eval(), no_grad, forward only - i.e. the native one-call
step.  The full model takes sentences of at most 126 tokens (C-Net's tiled convolution, include/umpr_b200.h): its L = 128 column runs at
L = 126 and says so; the review-net-only model (--workload music_small_r) runs the true L = 128.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="music_full")
ap.add_argument("--batches", default="64,256,1024,4096,8192")
ap.add_argument("--lengths", default="16,32,64,128")
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = syn.build_model(a.workload, syn.make_table(400003), seed=0, device=dev).eval()
from umpr_b200 import evaluate_mse  # noqa: E402
full = not syn.WORKLOADS[a.workload]["review_net_only"]
rows = []
for B in map(int, a.batches.split(",")):
    for L_req in map(int, a.lengths.split(",")):
        L = min(L_req, 126) if full else L_req
        b = syn.make_batch(a.workload, B, seed=B + L, L=L, skew=True)
        bd = (b[0].to(dev), b[1].to(dev), b[2].to(dev), b[3], b[4], b[5], b[6].to(dev), b[7].to(dev))
        with torch.no_grad():
            for _ in range(2):
                pred, _ = model(*bd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                pred, loss = model(*bd)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert bool(torch.isfinite(pred).all())
        mse = evaluate_mse(model, [bd])                  # the package-level evaluation loop on the same batch
        rows.append({"batch": B, "L": L_req, "L_run": L, "tokens": int(b[3].sum() + b[4].sum() + b[5].sum()), "ms": round(ms, 3),
                     "samples_per_s": round(B / ms * 1e3, 1), "us_per_token": round(ms * 1e3 / int(b[3].sum() + b[4].sum() + b[5].sum()), 4),
                     "max_valid_positions": int(max(b[3].sum(1).max(), b[4].sum(1).max())), "mse": round(mse, 4)})
        print(rows[-1], file=sys.stderr)
        del bd, pred
        torch.cuda.empty_cache()
print(json.dumps({"workload": a.workload, "mode": "eval forward (evaluate.py:6-14), length skew 90% 1-4 tokens / 10% full", "rows": rows}))
