"""Host-side cost of one train step (python + launches), to see when the step is host-bound.

    python profiles/prof_host.py [--batch 1024]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from umpr_b200 import synthetic as syn  # noqa: E402
from umpr_b200.train import FlatTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--workload", default="music_full")
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = syn.build_model(a.workload, syn.make_table(400003), seed=0, device=dev)
tr = FlatTrainer(model)
batches = []
for i in range(2):
    u, it, ui, ul, il, uil, ph, lab = syn.make_batch(a.workload, a.batch, seed=i)
    batches.append((u.to(dev), it.to(dev), ui.to(dev), ul, il, uil, ph.to(dev), lab.to(dev)))
for i in range(5):
    tr.train_step(batches[i % 2])
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for i in range(K):
    tr.train_step(batches[i % 2])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"batch {a.batch}: host issue time {1e3 * (t1 - t0) / K:.2f} ms/step, wall incl. GPU drain {1e3 * (t2 - t0) / K:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    tr.train_step(batches[i % 2])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(30)
