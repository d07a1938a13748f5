"""One profiled train step for ncu: warm up, then bracket exactly one step with cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python profiles/prof_step.py [--batch 1024] [--workload music_full]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import synthetic as syn  # noqa: E402
from umpr_b200.train import FlatTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--workload", default="music_full")
ap.add_argument("--vocab", type=int, default=400003)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = syn.build_model(a.workload, syn.make_table(a.vocab), seed=0, device=dev)
tr = FlatTrainer(model)
batches = []
for i in range(2):
    u, it, ui, ul, il, uil, ph, lab = syn.make_batch(a.workload, a.batch, vocab=a.vocab, seed=i)
    batches.append((u.to(dev), it.to(dev), ui.to(dev), ul, il, uil, ph.to(dev), lab.to(dev)))
for i in range(3):
    tr.train_step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.train_step(batches[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step, batch", a.batch)
