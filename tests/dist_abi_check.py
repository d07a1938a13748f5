"""2-rank check of the C-ABI NCCL helpers (launched by tests/test_gpu_parity.py::test_abi_allreduce_two_gpus via torchrun):
the library's communicator sums a flat bucket exactly like torch.distributed does, and a FlatTrainer step through it ends with
the same parameters as one through torch.distributed."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umpr_b200 import synthetic as syn  # noqa: E402
from umpr_b200.train import AbiComm, FlatTrainer, shard_batch  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
comm = AbiComm(rank, world, dev)
g = torch.Generator().manual_seed(5 + rank)
x = torch.randn(250_001, generator=g).to(dev)
a, b = x.clone(), x.clone()
comm.all_reduce(a)
dist.all_reduce(b)
torch.cuda.synchronize()
assert torch.equal(a, b), float((a - b).abs().max())
# one train step through either communicator
table = syn.make_table(3000, seed=2)
batch = syn.make_batch("music_full", 16, vocab=3000, seed=3)
shard = shard_batch(batch, rank, world)
out = []
for kind in ("torch", "abi"):
    model = syn.build_model("music_full", table, seed=1, device=dev)
    tr = FlatTrainer(model, lr=1e-3, comm=kind)
    tr.train_step(shard)
    torch.cuda.synchronize()
    out.append(tr.flat.clone())
# float atomics make two runs of the same step differ in the last bits: compare to 1e-6 of the parameter scale
err = float((out[0] - out[1]).abs().max() / out[0].abs().max())
assert err < 1e-6, err
# a short last batch: 1 sample over 2 ranks -> rank 1 gets no chunk (DataParallel would use one replica).  Both ranks must end with
# the parameters a single-GPU step on that one sample gives.
one = syn.make_batch("music_full", 1, vocab=3000, seed=4)
sh = shard_batch(one, rank, world)
assert (sh is None) == (rank == 1)
model = syn.build_model("music_full", table, seed=1, device=dev)
tr = FlatTrainer(model, lr=1e-3)
tr.train_step(sh)
solo = syn.build_model("music_full", table, seed=1, device=dev)
ts = FlatTrainer(solo, lr=1e-3, process_group=None)
ts.world = 1                                               # a private single-replica step on the same sample
ts.zero_grad()
pred, loss = solo(*one)
loss.backward()
ts.optimizer_step()
torch.cuda.synchronize()
err = float((tr.flat - ts.flat).abs().max() / ts.flat.abs().max())
assert err < 1e-6, ("short batch", rank, err)
comm.close()
dist.barrier()
if rank == 0:
    print("ABI_ALLREDUCE_OK")
dist.destroy_process_group()
