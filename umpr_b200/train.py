"""Train-step body (reference main.py:22-26,32-37) and its data-parallel form (main.py:81-82).

The reference wraps the model in ``torch.nn.DataParallel`` (single process, parameters re-broadcast every step,
gradients reduced onto GPU 0, Adam on GPU 0).  Here every GPU is its own process holding an identical replica:
all trainable parameters live in ONE flat fp32 buffer, their gradients in another; a step is

    flat_grad.zero_()  →  forward/backward on this rank's shard  →  all-reduce(flat_grad) over NCCL/NVLink
    →  fused Adam(+L2 on non-bias tensors) with the 1/world scale folded in  (same update on every rank).

``loss.mean()`` over DataParallel's k replica losses (main.py:34) weights each shard's mean loss by 1/k, which is
exactly ``sum of shard gradients / k``.  Shards are the ``torch.chunk`` pieces DataParallel's scatter would produce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr


COMM_CTAS = 4        # CTAs NCCL may use for the ~1 MB gradient all-reduce (NCCL_MAX_CTAS, set before the communicator is created)


class AbiComm:
    """The library's NCCL communicator (C-ABI ``umpr_comm_*``): rank 0 creates the unique id, ``torch.distributed`` carries its 128
    bytes to the other ranks (any backend), every rank joins; ``all_reduce`` sums a flat fp32 CUDA tensor in place on the current stream."""

    def __init__(self, rank, world, device, process_group=None):
        import ctypes
        import os
        os.environ.setdefault("NCCL_MAX_CTAS", str(COMM_CTAS))      # latency-bound message: a few CTAs are enough, and they must fit
        lib = _lib.load()                                            # beside the persistent GRU backward (plan.SMS_RESERVED_FOR_COMM)
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            if lib.umpr_comm_unique_id(raw) != 0:
                raise RuntimeError(f"umpr_comm_unique_id: {_lib.last_error()}")
            buf = torch.tensor(list(raw), dtype=torch.uint8)
        backend = dist.get_backend(process_group)
        t = buf.to(device) if backend == "nccl" else buf
        dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
        ident = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
        self._h = ctypes.c_void_p()
        torch.cuda.set_device(device)
        if lib.umpr_comm_init(rank, world, ident, ctypes.byref(self._h)) != 0:
            raise RuntimeError(f"umpr_comm_init: {_lib.last_error()}")
        self.rank, self.world = rank, world

    def all_reduce(self, flat):
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        call("umpr_allreduce", self._h, ptr(flat), flat.numel())

    def close(self):
        if self._h:
            _lib.load().umpr_comm_destroy(self._h)
            self._h = None


def pin_rank_to_cores(local_rank: int, local_world: int):
    """Give every rank of a node its own slice of the host cores (one process per GPU): the step is issued by one Python thread next to
    the plan-prefetch workers, and a rank whose threads get migrated - or share a physical core with another rank's - stalls all ranks
    at the all-reduce.  Logical CPUs are grouped by physical core (hyper-thread siblings stay together) before they are dealt out."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    groups, seen = [], set()
    for c in cpus:
        if c in seen:
            continue
        sib = {c}
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                for part in f.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    sib |= set(range(int(lo), int(hi or lo) + 1))
        except (OSError, ValueError):
            pass
        g = sorted(sib & set(cpus))
        seen |= set(g)
        groups.append(g)
    ordered = [c for g in groups for c in g]             # siblings adjacent: a slice takes whole physical cores first
    per = len(ordered) // max(1, local_world)
    if per < 1:
        return None
    mine = ordered[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    # one intra-op thread: the host work of a rank is many small calls (a 20 k-element sort, a few numpy passes) on the issuing
    # thread and the plan workers - extra OpenMP threads per call would only fight them for this rank's few cores
    torch.set_num_threads(1)
    # the issuing thread gets the first physical core of the slice to itself (pin_issuing_thread), the plan workers the rest
    global _ISSUE_CPUS, _WORKER_CPUS
    first = [c for c in groups[[g[0] for g in groups].index(mine[0])] if c in mine] if mine else []
    rest = [c for c in mine if c not in first]
    _ISSUE_CPUS, _WORKER_CPUS = (first, rest) if first and rest else (None, None)
    return mine


_ISSUE_CPUS = None
_WORKER_CPUS = None


def pin_issuing_thread():
    """Call from the thread that issues the steps, AFTER the communicators exist (their helper threads inherit the caller's mask at
    creation and should keep the whole slice): from now on this thread owns one physical core and the plan-prefetch workers run on
    the others.  With eight ranks on a 32-thread host a rank has two physical cores; the reference's own ``torch.sort`` calls on the
    workers (~2 ms of CPU per 3.4 ms step) otherwise share a core with the thread whose launches must stay a step ahead of the GPU -
    the branches of a step only overlap on the GPU when all of them are already queued."""
    import os
    if _ISSUE_CPUS:
        os.sched_setaffinity(0, _ISSUE_CPUS)
    return _ISSUE_CPUS


def _pin_worker_thread():
    import os
    if _WORKER_CPUS:
        try:
            os.sched_setaffinity(0, _WORKER_CPUS)
        except OSError:
            pass


class FlatTrainer:
    def __init__(self, model, lr=1e-6, weight_decay=1e-3, betas=(0.9, 0.999), eps=1e-8, lr_decay=0.99, process_group=None, comm=None,
                 native=True):
        """``comm``: "torch" = ``torch.distributed.all_reduce`` on ``process_group`` (default), "abi" = the library's own NCCL
        communicator (``umpr_comm_init`` / ``umpr_allreduce``, include/umpr_b200.h), bootstrapped by broadcasting the NCCL unique
        id through ``torch.distributed``; default from ``UMPR_COMM``."""
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise RuntimeError("nothing to train")
        # bucket layout [early | late]: R-Net's GRU tensors go last - their gradients are the last to become final (the fused R-Net GRU
        # backward is the last kernel of a step), everything else can be all-reduced while that kernel runs
        late = lambda n: n.startswith("review_net.r_net.gru.")
        named = [x for x in named if not late(x[0])] + [x for x in named if late(x[0])]
        dev = named[0][1].device
        ALIGN = 64                                                  # floats: every tensor starts on a 256-byte boundary
        up = lambda k: (k + ALIGN - 1) // ALIGN * ALIGN             # (the kernels use 128-bit loads on some weights)
        total = sum(up(p.numel()) for _, p in named)
        self.model, self.names = model, [n for n, _ in named]
        self.n_params = sum(p.numel() for _, p in named)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # the gradient bucket carries one extra element: 1.0 on every rank that received a chunk of the batch.  After the all-reduce
        # it holds the number of non-empty shards, which is what DataParallel's loss.mean() divides by (main.py:34)
        self.bucket = torch.zeros(total + ALIGN, dtype=torch.float32, device=dev)
        self.grad = self.bucket[:total]
        self.shard_count = self.bucket[total:total + 1]
        self.wd = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        o = 0
        self.n_early = total
        with torch.no_grad():
            for n, p in named:
                k = p.numel()
                if late(n) and self.n_early == total:
                    self.n_early = o
                self.flat[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat[o:o + k].view_as(p)          # parameters become views of the flat buffer
                p.grad = self.grad[o:o + k].view_as(p)          # autograd accumulates straight into the bucket
                self.wd[o:o + k] = 0.0 if "bias" in n else weight_decay      # main.py:23-24
                o += up(k)
        self._grad_ptrs = [(n, p, p.grad.data_ptr()) for n, p in named]
        self.lr, self.betas, self.eps, self.lr_decay = lr, betas, eps, lr_decay
        self.step_no = 0
        self.native_steps = 0          # steps taken through the native one-call path
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.comm = None
        import os
        # the gradient exchange goes through the library's own NCCL communicator (C-ABI) by default; "torch" = torch.distributed.all_reduce
        if (comm or os.environ.get("UMPR_COMM", "abi")) == "abi" and self.world > 1 and self.flat.is_cuda:
            self.comm = AbiComm(dist.get_rank(process_group), self.world, dev, process_group)
        # the whole forward + backward as one native call (csrc/step.cu) whenever the model is the standard UMPR on a CUDA device
        # and the batch lies inside that path's envelope; the autograd path (functional.py) is the general fallback
        self.native = None
        if native and self.flat.is_cuda and os.environ.get("UMPR_NATIVE_STEP", "1") == "1":
            from .model import UMPR
            from .step import NativeStep
            if isinstance(model, UMPR):
                self.native = NativeStep(model, with_grads=True)
        # UMPR_OVERLAP=1: two buckets all-reduced from inside umpr_step (csrc/step.cu, umpr_step_comm), the first on a communication
        # stream while the R-Net GRU backward runs.  Off by default since the step runs its branches on several streams: the C-Net
        # branch now ends together with the R-Net GRU backward, so the first bucket is final only at the very end, and NCCL's CTAs wait
        # for SMs that the persistent kernels of two branches hold (2 GPUs, batch 1024: 3.80 ms per step with, 3.63 ms without; no
        # exchange at all 3.61 ms - the ~1 MB all-reduce after the backward costs ~0.03 ms).
        self.overlap = self.native is not None and self.comm is not None and os.environ.get("UMPR_OVERLAP", "0") == "1"
        self._skip_reduce = os.environ.get("UMPR_DIAG_NO_REDUCE", "0") == "1"     # diagnostics only: ranks run unsynchronised (wrong training)
        if self._skip_reduce:
            self.overlap = False
        if self.overlap:
            from . import plan as plan_mod
            plan_mod.SMS_RESERVED_FOR_COMM = COMM_CTAS          # the R-Net GRU launches leave these SMs to the all-reduce kernel

    def zero_grad(self):
        self.bucket.zero_()

    def check_bucket(self):
        """The kernels (and autograd) accumulate into ``p.grad``, which must still be the view into the flat bucket: ``model.zero_grad()``
        (set_to_none) or an outside optimizer detaches it, and Adam would then silently step on zeros."""
        for n, p, addr in self._grad_ptrs:
            if p.grad is None or p.grad.data_ptr() != addr:
                raise RuntimeError(f"umpr_b200: the gradient of {n!r} no longer lives in FlatTrainer's bucket (was model.zero_grad() or "
                                   "another optimizer used?) - use FlatTrainer.zero_grad()")

    def reduce_gradients(self):
        """The flat bucket (0.57–1.0 MB, latency-bound): one all-reduce through the library's communicator - or, with UMPR_OVERLAP=1, the
        two collectives [early | late] in the order the native step issues them from inside its backward."""
        if self.comm is not None:
            if self.overlap:          # a rank without a shard joins the two collectives the native step issues (same order)
                self.comm.all_reduce(self.bucket[:self.n_early])
                self.comm.all_reduce(self.bucket[self.n_early:])
            else:
                self.comm.all_reduce(self.bucket)
        elif self.world > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self, use_shard_count: bool = False):
        """``use_shard_count``: scale by 1 / (number of ranks that had a shard), read on the device from the all-reduced bucket."""
        self.step_no += 1
        if self.flat.is_cuda:
            call("umpr_adam_step", ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), ptr(self.wd), self.flat.numel(),
                 float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_no, 1.0 / self.world,
                 ptr(self.shard_count) if use_shard_count else None)
        else:
            raise RuntimeError("umpr_b200: optimizer runs on the GPU only")

    def end_epoch(self):
        self.lr *= self.lr_decay                                   # ExponentialLR, main.py:26,54

    def train_step(self, batch):
        """main.py:32-37 on this rank's shard.  Returns (prediction, loss) of the shard.  ``batch=None``: this rank got no chunk of
        a short last batch (``shard_batch``) - it still takes part in the all-reduce and applies the same update as everybody else."""
        from . import functional as F
        if not self.model.training:          # nn.Module.train() walks the whole module tree: ~0.2 ms of host time per step when repeated
            self.model.train()
        self.zero_grad()
        pred = loss = None
        reduced = False
        if batch is not None:
            self.check_bucket()
            if self.world > 1:
                self.shard_count.fill_(1.0)         # this rank contributes a shard (counted by the all-reduce, read by the Adam kernel)
            plans = self.native.plans_of(batch, self.flat.device) if self.native is not None else None
            if plans is not None and self.native.supported(batch, plans):
                # one C-ABI call: forward, backward, gradients accumulated into the flat bucket - and, with the library's communicator,
                # both all-reduces issued from inside it so that the first overlaps the last backward kernel
                with torch.cuda.device(self.flat.device):
                    if self.overlap:
                        self._set_overlap(True)
                    try:
                        pred, loss = self.native.run(batch, True, plans, routing_log=F.ROUTING_LOG)
                    finally:
                        if self.overlap:
                            self._set_overlap(False)
                reduced = self.overlap
                self.native_steps += 1
            else:
                pred, loss = self.model(*batch)
                F.DIRECT_GRAD_ACCUM = True      # the kernels accumulate parameter gradients straight into the flat bucket (functional._sinks)
                try:
                    (loss if loss.dim() == 0 else loss.mean()).backward()      # main.py:34 (the mean over replica losses is the identity for one shard)
                finally:
                    F.DIRECT_GRAD_ACCUM = False
        if not reduced and not self._skip_reduce:
            self.reduce_gradients()
        self.optimizer_step(use_shard_count=self.world > 1)
        return pred, loss

    def _set_overlap(self, on: bool):
        lib = _lib.load()
        if lib.umpr_step_comm(self.comm._h if on else None, ptr(self.bucket), self.n_early, self.bucket.numel()) != 0:
            raise RuntimeError(f"umpr_step_comm: {_lib.last_error()}")


def prepare_batch(batch, device):
    """Host-side half of ImprovedRnn's packing for one batch (the reference's ``torch.sort`` call and the integer plan derived
    from it, plan.PackPlan) - what a collate worker would do.  The plans ride on the three ``lengths``
    tensors of the 8-tuple, so the model's forward signature is unchanged; no CUDA call is made here."""
    from .plan import PackPlan
    u, it, ui, ul, il, uil = batch[:6]
    for ids, lens in ((u, ul), (it, il), (ui, uil)):
        if lens.numel() and ids.dim() == 3:
            lens._umpr_plan = PackPlan(lens.reshape(-1), ids.shape[2], device, upload=False)
            if lens._umpr_plan.R == 128 and ids.shape[2] <= 128:
                lens._umpr_plan._snet_host()                      # S-Net / C-Net tile tables (numpy; uploaded by the consumer)
                if ids.shape[2] <= 126:
                    lens._umpr_plan._cnet_host()
    # tile schedules of the two fused GRU launches (user+item; ui+user+item), for the native one-call step
    pl = [getattr(t, "_umpr_plan", None) for t in (ul, il, uil)]
    if pl[0] is not None and pl[1] is not None and pl[0].R == 128 and pl[1].R == 128:
        from .plan import build_schedule
        from . import plan as plan_mod
        n_sm = _lib.sm_count(device) if torch.device(device).type == "cuda" else 148
        n_ctas = max(1, n_sm // 2)
        sched = [build_schedule([pl[0].tile_len, pl[1].tile_len], max(1, (n_sm - plan_mod.SMS_RESERVED_FOR_COMM) // 2))]
        if pl[2] is not None and pl[2].R == 128:
            sched.append(build_schedule([pl[2].tile_len, pl[0].tile_len, pl[1].tile_len], n_ctas))
        ul._umpr_sched = sched
    return batch


class PlanPrefetcher:
    """Iterator over batches that prepares the NEXT batches' pack plans on worker threads while the current step runs (the
    data-loader pattern; ``torch.sort`` and the C plan builders release the GIL).  ``for batch in PlanPrefetcher(loader, device): ...``
    Batches come out in the loader's order.  The reference's own ``torch.sort`` call is most of a plan's cost (~0.7 ms per review
    side at batch 1024), so two workers keep up with a 5 ms step even when eight ranks share one host."""

    def __init__(self, batches, device, depth: int = 3, workers: int = 2):
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        import sys
        self._q = queue.Queue(maxsize=depth)
        self._done = object()
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="umpr-plan", initializer=_pin_worker_thread)
        # the step is issued by ONE Python thread: with CPython's default 5 ms switch interval a worker in a pure-Python stretch can
        # keep the GIL for as long as a whole step takes.  0.2 ms bounds what the issuing thread can lose to the workers.
        if sys.getswitchinterval() > 2e-4:
            sys.setswitchinterval(2e-4)

        def feed():                             # pulls from the loader and hands the batches to the pool, in order
            try:
                for b in batches:
                    self._q.put(self._pool.submit(prepare_batch, b, device))
            except BaseException as e:          # surfaced on the consumer's thread
                self._q.put(e)
            self._q.put(self._done)

        def feed_pinned():
            _pin_worker_thread()
            feed()

        self._thr = threading.Thread(target=feed_pinned, daemon=True)
        self._thr.start()

    def __iter__(self):
        return self

    def __next__(self):
        f = self._q.get()
        if f is self._done:
            self._pool.shutdown(wait=False)
            raise StopIteration
        if isinstance(f, BaseException):
            raise f
        return f.result()


class AsyncScalarReader:
    """Per-step device→host read of a scalar (the loss of main.py:38-39) that does not drain the GPU: the value goes to a
    pinned slot with an asynchronous copy, and is handed out one step later, when its copy has completed.  Every step's value
    is still delivered, in order; ``drain()`` returns the ones still in flight."""

    def __init__(self, depth: int = 2):
        self.depth = depth
        self.slots = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.head = 0          # next slot to write
        self.inflight = 0

    def push(self, scalar: torch.Tensor):
        """Queue the read of ``scalar``; returns the list of values whose reads had to complete to make room (0 or 1)."""
        out = []
        if self.inflight == self.depth:
            out.append(self._pop())
        i = self.head
        self.slots[i].copy_(scalar.detach().reshape(1), non_blocking=True)
        self.events[i].record()
        self.head = (i + 1) % self.depth
        self.inflight += 1
        return out

    def _pop(self):
        i = (self.head - self.inflight) % self.depth
        self.events[i].synchronize()
        self.inflight -= 1
        return float(self.slots[i])

    def drain(self):
        return [self._pop() for _ in range(self.inflight)]


def shard_batch(batch, rank: int, world: int):
    """The chunk ``torch.nn.DataParallel``'s scatter hands to replica ``rank`` (``torch.chunk`` along dim 0).
    Returns None when there are fewer chunks than ranks (e.g. B=9, world=8 → 5 chunks)."""
    out = []
    for t in batch:
        if t.dim() == 0 or t.numel() == 0:
            out.append(t)
            continue
        chunks = torch.chunk(t, world, dim=0)
        if rank >= len(chunks):
            return None
        out.append(chunks[rank])
    return tuple(out)
