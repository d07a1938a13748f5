"""Data-parallel host logic on CPU with gloo, world_size 2:
shards are DataParallel's torch.chunk pieces, each rank sorts its own lengths, total_length stays the global L, and the
all-reduced flat gradient bucket divided by the number of shards equals the mean over shards of the oracle's gradients
(main.py:34 ``loss.mean()`` over replica losses)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import umpr_oracle as orc
from umpr_b200.plan import PackPlan
from umpr_b200.train import shard_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    c = cases.CASES["umpr_full_v4"]
    params = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    batch = cases.make_batch(c)
    shard = shard_batch(batch, rank, world)
    # per-shard plan: own sort, global total_length
    plan = PackPlan(shard[3], batch[0].shape[2], "cpu", tile_rows=32)
    assert plan.L == batch[0].shape[2] and plan.N == shard[3].numel()
    _, loss, grads = orc.umpr_loss_and_grads(params, shard, review_net_only=c["review_net_only"])
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in keys])          # the flat bucket
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                     # FlatTrainer.reduce_gradients
    flat /= world                                                   # folded into the Adam kernel as grad_scale = 1/world
    if rank == 0:
        torch.save({"flat": flat, "keys": keys, "loss": loss}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_bucket_matches_mean_of_shard_gradients(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    c = cases.CASES["umpr_full_v4"]
    params = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    batch = cases.make_batch(c)
    ref = None
    for r in range(2):
        _, _, g = orc.umpr_loss_and_grads(params, shard_batch(batch, r, 2), review_net_only=c["review_net_only"])
        f = torch.cat([g[k].reshape(-1) for k in got["keys"]])
        ref = f if ref is None else ref + f
    ref /= 2
    assert torch.allclose(got["flat"], ref, rtol=1e-5, atol=1e-8)


def test_shards_follow_dataparallel_chunking():
    t = torch.arange(37 * 3).reshape(37, 3)
    sizes = [shard_batch((t,), r, 5)[0].shape[0] for r in range(5)]
    assert sizes == [8, 8, 8, 8, 5]                                  # torch.chunk: ceil split
    nine = torch.arange(9)
    assert [None if shard_batch((nine,), r, 8) is None else shard_batch((nine,), r, 8)[0].numel() for r in range(8)] == [2, 2, 2, 2, 1, None, None, None]
    # empty photos tensor (review_net_only, dataset.py:158,180) passes through untouched
    assert shard_batch((torch.zeros(0), nine), 1, 2)[0].numel() == 0


def test_flat_trainer_layout_cpu_tensors():
    """FlatTrainer's bucket: every parameter is a 256-byte aligned view; bias elements carry no weight decay (main.py:23-24)."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import FlatTrainer
    m = syn.build_model("music_small_r", syn.make_table(50), device="cpu")
    ref = {k: v.detach().clone() for k, v in m.named_parameters() if v.requires_grad}
    tr = FlatTrainer(m, lr=1e-3, weight_decay=1e-3)
    assert tr.n_params == 143105
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        assert torch.equal(p.detach(), ref[k])
        assert p.data_ptr() % 256 == tr.flat.data_ptr() % 256
        off = (p.data_ptr() - tr.flat.data_ptr()) // 4
        wd = tr.wd[off:off + p.numel()]
        assert float(wd.min()) == float(wd.max()) and abs(float(wd.min()) - (0.0 if "bias" in k else 1e-3)) < 1e-9
        assert p.grad.data_ptr() == tr.grad.data_ptr() + off * 4
    with pytest.raises(RuntimeError, match="GPU only"):
        tr.optimizer_step()


def test_rank_pinning_gives_the_issuing_thread_its_own_cores():
    """``pin_rank_to_cores`` deals the host's CPUs out to the ranks of a node (disjoint slices); ``pin_issuing_thread`` then keeps one
    physical core of the slice for the calling thread and leaves the others to the plan workers.  Run in a child process: affinity is
    process state."""
    import subprocess
    import sys
    code = (
        "import os, json, sys\n"
        "sys.path.insert(0, %r)\n"
        "from umpr_b200 import train\n"
        "all_cpus = sorted(os.sched_getaffinity(0))\n"
        "out = []\n"
        "for r in range(2):\n"
        "    os.sched_setaffinity(0, all_cpus)\n"
        "    mine = train.pin_rank_to_cores(r, 2)\n"
        "    issue = train.pin_issuing_thread()\n"
        "    out.append([mine, issue, train._WORKER_CPUS, sorted(os.sched_getaffinity(0))])\n"
        "print(json.dumps([all_cpus, out]))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import json
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    all_cpus, out = json.loads(res.stdout.strip().splitlines()[-1])
    if len(all_cpus) < 4:
        pytest.skip("needs at least 4 host CPUs")
    (m0, i0, w0, a0), (m1, i1, w1, a1) = out
    assert m0 and m1 and not set(m0) & set(m1) and set(m0) | set(m1) <= set(all_cpus)
    for mine, issue, workers, now in out:
        if len(mine) >= 2:
            assert issue and workers and not set(issue) & set(workers) and set(issue) | set(workers) == set(mine)
            assert now == sorted(issue)          # the calling thread sits on the issuing cores only
