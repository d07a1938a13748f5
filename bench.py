#!/usr/bin/env python
"""bench.py — UMPR train-step samples/sec on N B200s (one process per GPU), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--workload music_full]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" = main.py:32-37 on one synthetic batch: zero_grad, forward, backward, gradient all-reduce (N>1), fused Adam.
``value``: inputs already resident in HBM.  ``e2e``: the same step through the public API (``UMPR.forward`` of the
drop-in module) with HOST (pinned) input buffers, the H2D copies and the D2H read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "UMPR train samples/sec (device-timed, fwd+bwd+allreduce+Adam)"
UNIT = "samples/s"

# which roofline bounds each entry point: dense contractions -> tensor pipe, the rest -> HBM
TENSOR_BOUND = {"umpr_gru_inproj", "umpr_gru_inproj_tc", "umpr_gru_recurrence_fwd", "umpr_gru_recurrence_bwd", "umpr_gru_wgrad",
                "umpr_gru_wgrad_tc", "umpr_gru_fwd_tc", "umpr_gru_bwd_tc", "umpr_sgemm", "umpr_tc_gemm_ws",
                "umpr_tc_gemm_nt", "umpr_coattn_fwd", "umpr_coattn_fwd_tc", "umpr_snet_fwd", "umpr_snet_bwd", "umpr_cnet_conv_fwd",
                "umpr_cnet_conv_fwd_tc", "umpr_snet_fwd_tc", "umpr_snet_bwd_tc", "umpr_tc_gemm_tn"}
# arithmetic each entry point runs in (everything is fp32 in and out; "3xBF16" = fp32 operands split into bf16 hi+lo, fp32 accumulate)
MATH = {"umpr_gru_fwd_tc": "tcgen05 kind::f16, 3xBF16 split, fp32 accumulation in TMEM; gates packed fp32x2 (ex2/rcp approx)",
        "umpr_gru_bwd_tc": "tcgen05 3xBF16: gates RECOMPUTED from the operand images (not counted as algorithmic work), carry product and weight gradients (accumulated in TMEM); element-wise packed fp32x2",
        "umpr_tc_gemm_ws": "tcgen05 3xBF16", "umpr_tc_gemm_nt": "tcgen05 3xBF16", "umpr_cnet_conv_fwd_tc": "tcgen05 3xBF16 + exact fp32 re-scoring of near-ties",
        "umpr_coattn_fwd_tc": "tcgen05 3xBF16 over the valid rows only + exact fp32 re-scoring of near-ties",
        "umpr_snet_fwd_tc": "tcgen05 3xBF16 over the valid rows only; tanh / softmax fp32",
        "umpr_snet_bwd_tc": "tcgen05 3xBF16 (scores recomputed; dx and dMs^T products, dMs accumulated in TMEM); element-wise fp32",
        "umpr_tc_gemm_tn": "tcgen05 3xBF16, both operands MN-major"}
# the kernels BASELINE.json's north_star names: reported next to the dominant one
NAMED = ["umpr_gru_fwd_tc", "umpr_gru_bwd_tc", "umpr_gather_pack_tc", "umpr_gather_pack", "umpr_coattn_fwd_tc", "umpr_coattn_fwd", "umpr_coattn_bwd",
         "umpr_snet_fwd_tc", "umpr_snet_bwd_tc", "umpr_snet_fwd", "umpr_snet_bwd"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled during the timed region (B200_PROFILING.md 'clocks' line)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            if os.environ.get("UMPR_BENCH_NO_CLOCKS") == "1":      # diagnostics: no NVML polling during the timed region
                raise RuntimeError
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        try:
            from umpr_b200.train import _pin_worker_thread
            _pin_worker_thread()            # not on the issuing thread's core
        except Exception:
            pass
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            # ~6 samples over the ~70 ms timed region of the default run; polling faster costs throughput (NVML takes driver locks:
            # 2 GPUs, 4 ms interval: 3.62 ms per step against 3.49 ms without any polling)
            self._stop.wait(0.012)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently turn the
    CPU arm into a single-thread run: the count is taken from the affinity mask and set explicitly."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return n


def load_reference():
    """The reference's own ``src/model.py``, unmodified, from the git-ignored staging directory baseline/_ref/ (written by
    ``__graft_entry__.build()`` in the build container; it travels to the GPU box with the snapshot).  Only harness patch
    (SURVEY.md §8c): torchvision's VGG16 constructor - which would download weights - becomes ``nn.Flatten``, so ``photos`` carries
    the backbone's 1000-d output features and model.py:216-218 passes them on unchanged.  → module, or None when not staged."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "src", "model.py")):
        return None
    import importlib
    import torchvision
    torchvision.models.vgg16 = lambda pretrained=True, num_classes=1000: torch.nn.Flatten()
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    return importlib.import_module("src.model")


def cpu_reference_run(workload, batch, steps, warmup, seed=0):
    """main.py:22-25,32-37 with the reference's own modules on the host cores: UMPR(config, word_emb) from baseline/_ref,
    the weights of the GPU arm, torch.optim.Adam with the reference's parameter groups, forward + backward + step.
    → (samples/s, ms/step, kind)."""
    from umpr_b200 import synthetic as syn
    ref = load_reference()
    if ref is None:
        v, ms = cpu_port_run(workload, batch, steps, warmup, seed)
        return v, ms, "port"
    table = syn.make_table(400003, seed=0)
    ours = syn.build_model(workload, table, seed=0, device="cpu")           # parameter container only: same init as the GPU arm
    model = ref.UMPR(syn.workload_config(workload), table.numpy())
    r = model.load_state_dict(ours.state_dict(), strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    del ours
    opt = torch.optim.Adam([
        {'params': (p for name, p in model.named_parameters() if 'bias' not in name)},
        {'params': (p for name, p in model.named_parameters() if 'bias' in name), 'weight_decay': 0.}
    ], 1e-6, weight_decay=1e-3)                                              # main.py:22-25 (config.py:13-14)
    rno = syn.WORKLOADS[workload]["review_net_only"]
    batches = []
    for i in range(2):
        b = list(syn.make_batch(workload, batch, seed=seed + i))
        if not rno:
            b[6] = b[6].reshape(*b[6].shape, 1, 1)                           # features where the reference expects images
        batches.append(b)
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        model.train()                                                        # main.py:32-37
        pred, loss = model(*batches[i % 2])
        loss = loss.mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, "reference"


def cpu_port_run(workload, batch, steps, warmup, seed=0):
    """Fallback when baseline/_ref is not staged: the oracle port with the library GRU calls of model.py:18-20,
    forward + backward + torch.optim.Adam (main.py:22-25,32-37), all host threads."""
    from oracle import umpr_oracle as orc     # the one place bench.py executes oracle/: the CPU baseline
    from umpr_b200 import synthetic as syn
    table = syn.make_table(400003, seed=0)
    model = syn.build_model(workload, table, seed=0, device="cpu")
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in params.items():
        v.requires_grad_(k != "embedding.weight")
    named = [(k, v) for k, v in params.items() if v.requires_grad]
    opt = torch.optim.Adam([{"params": [v for k, v in named if "bias" not in k]},
                            {"params": [v for k, v in named if "bias" in k], "weight_decay": 0.0}], 1e-6, weight_decay=1e-3)
    rno = syn.WORKLOADS[workload]["review_net_only"]
    batches = [syn.make_batch(workload, batch, seed=seed + i) for i in range(2)]
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        pred, loss = orc.umpr_forward(params, batches[i % 2], review_net_only=rno, impl="lib")
        opt.zero_grad()
        loss.mean().backward()
        opt.step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def workload_name(workload):
    return (f"{workload}: full UMPR, Amazon Digital Music shape (S=20, L=20, S_ui=5, V=1, Pc=1, GloVe-50d table 400003x50, VGG16 features)"
            if workload == "music_full" else workload)


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the same step on the box's host cores, on the GPU arm's workload and
    per-GPU batch.  Under torchrun only rank 0 works; every step is one per-GPU batch (a bounded sample of the N-GPU global batch:
    CPU throughput in samples/s does not depend on how many such batches make up a global step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    B = args.ref_batch or args.batch
    val, ms, kind = cpu_reference_run(args.workload, B, args.steps, args.warmup)
    what = "unmodified src/model.py from baseline/_ref (VGG16 -> Flatten: photos are features)" if kind == "reference" else "oracle port with torch's CPU GRU"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "per_gpu_batch": B, "global_batch": B * args.gpus,
                       "parallelism": f"dp{args.gpus}", "step": "zero_grad+fwd+bwd+adam (main.py:32-37)", "device": "cpu",
                       "host_threads": cores, "os_cpu_count": os.cpu_count(),
                       "sample": f"each step = one batch of {B} samples on the host cores" + (f" (1/{args.gpus} of the global batch)" if args.gpus > 1 else "")},
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} train steps of batch {B} ({args.workload}), {what}, fwd+bwd+Adam"},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def tensor_bytes(ts):
    return int(sum(t.numel() * t.element_size() for t in ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="umpr_b200", choices=["umpr_b200", "reference"])
    ap.add_argument("--workload", default="music_full")
    ap.add_argument("--batch", type=int, default=1024, help="per-GPU batch (weak scaling)")
    ap.add_argument("--ref-batch", type=int, default=0, help="batch of one CPU reference step (default: the GPU arm's per-GPU batch)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch per GPU; strong: --batch is the GLOBAL batch, split over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch64", action="store_true")
    ap.add_argument("--kernel-table", action="store_true", help="print the per-entry-point time table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import umpr_b200
    from umpr_b200 import _lib, synthetic as syn
    from umpr_b200.train import FlatTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cores = None
    if world > 1:
        from umpr_b200.train import pin_rank_to_cores
        cores = pin_rank_to_cores(local, world)         # one process per GPU, each on its own host cores
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    umpr_b200.require_lib()
    peaks = load_peaks()
    B, K, W = args.batch, args.steps, args.warmup
    if args.scaling == "strong":          # fixed GLOBAL batch: every rank takes the chunk DataParallel's scatter would give it
        if B % world:
            raise SystemExit(f"bench.py: --scaling strong needs a global batch divisible by the world size ({B} % {world})")
        B = B // world

    table = syn.make_table(400003, seed=0)
    model = syn.build_model(args.workload, table, seed=0, device=dev)       # identical replicas: same seed on every rank
    trainer = FlatTrainer(model, lr=1e-6, weight_decay=1e-3)
    n_params = trainer.n_params
    if world > 1:
        from umpr_b200.train import pin_issuing_thread
        issue_cpus = pin_issuing_thread()              # (after the communicators exist: their threads keep the rank's whole slice)
    else:
        issue_cpus = None
    NB = 4
    host = [syn.make_batch(args.workload, B, seed=1000 * rank + i) for i in range(NB)]
    pin = lambda t: t.pin_memory() if t.numel() else t
    host = [tuple(pin(t) for t in b) for b in host]

    def resident(b):     # reviews / photos / labels in HBM; lengths stay on the host as the reference's collate leaves them
        u, it, ui, ul, il, uil, ph, lab = b
        return (u.to(dev), it.to(dev), ui.to(dev), ul, il, uil, ph.to(dev), lab.to(dev))

    devb = [resident(b) for b in host]
    tokens = [int(b[3].sum() + b[4].sum() + b[5].sum()) for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = {}

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            # every rank's own device time and SM clock (a slow or throttled GPU shows up here); the reported time is the MAX over ranks
            mine = torch.tensor([float(ms), float((clocks or {}).get("sm_mhz") or 0)], device=dev)
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            per_rank["ms_per_step"] = [round(float(t[0]) / steps, 4) for t in every]
            per_rank["sm_mhz"] = [int(t[1]) for t in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms), clocks

    from umpr_b200.train import PlanPrefetcher

    def stream_of(batches, n):            # what a data loader does: the next batch's host-side pack plans are prepared on a worker
        def gen():                        # thread while the current step runs (fresh plans every step: nothing is cached across steps)
            for i in range(n):
                b = batches[i % NB]
                yield tuple(t.clone() if j in (3, 4, 5) else t for j, t in enumerate(b))     # new lengths tensors, as a loader would emit
        return PlanPrefetcher(gen(), dev)

    feed = {"it": None}

    host_t = {"wait_batch": 0.0, "issue": 0.0, "n": 0}

    def step_resident(i):
        t0 = time.perf_counter()
        b = next(feed["it"])
        t1 = time.perf_counter()
        trainer.train_step(b)
        host_t["wait_batch"] += t1 - t0                  # host time waiting for the prefetched batch (plans)
        host_t["issue"] += time.perf_counter() - t1      # host time inside train_step (asynchronous launches)
        host_t["n"] += 1

    sink = []
    from umpr_b200.train import AsyncScalarReader
    reader = AsyncScalarReader(depth=2)

    def step_e2e(i):
        pred, loss = trainer.train_step(next(feed["it"]))      # H2D of ids/photos/labels happens inside UMPR.forward
        sink.extend(reader.push(loss))                         # main.py:39: D2H read of the loss EVERY step (pinned slot, async copy,
                                                               # delivered one step later so the host keeps issuing the next step)

    # ---- warm-up, with every entry point timed once to find the dominant kernel.  The step runs through ONE native call
    # (umpr_step, csrc/step.cu); its entry points are timed by CUDA-event pairs recorded inside that call on the launching stream
    native = trainer.native is not None
    NS = type(trainer.native) if native else None
    feed["it"] = stream_of(devb, W)
    for i in range(W - 1):
        step_resident(i)
    probe = next(feed["it"])
    if native:
        plans = NS.plans_of(probe, dev)
        native = trainer.native.supported(probe, plans)
    if native:
        work = trainer.native.work_table(probe, plans)
        NS.serialize(True)            # per-kernel times without co-scheduled kernels: this one step runs on a single stream
        NS.profile_begin(None)
        n0 = trainer.native_steps
        trainer.train_step(probe)
        table_ms = NS.profile_end()
        NS.serialize(False)
        assert trainer.native_steps == n0 + 1
        for k, v in table_ms.items():
            v["flops"], v["bytes"] = work.get(k, [0.0, 0.0])
    else:
        _lib.start_timing()
        trainer.train_step(probe)
        table_ms = _lib.stop_timing()
    top = max(table_ms, key=lambda k: table_ms[k]["ms"])
    step_ms_profiled = sum(v["ms"] for v in table_ms.values())
    if args.kernel_table and rank == 0:
        for k, v in sorted(table_ms.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"  {k:28s} calls {v['calls']:3d}  {v['ms']:9.3f} ms  {100 * v['ms'] / step_ms_profiled:5.1f}%  "
                  f"{v['flops'] / 1e9:10.2f} GFLOP {v['bytes'] / 1e6:10.1f} MB", file=sys.stderr)

    # ---- timed region: K steps, inputs resident; only the dominant entry point carries event pairs
    launches0 = _lib.launch_count
    if not native:
        _lib.start_timing(only=[top])
    feed["it"] = stream_of(devb, K)
    n0 = trainer.native_steps
    host_t.update(wait_batch=0.0, issue=0.0, n=0)
    # clocks and throttle reasons during the timed region: polled by rank 0 only - NVML calls take driver locks that every process on
    # the box feels (8 ranks polling every 12 ms: 4.31 ms per step against 3.94 ms in the region without polling)
    ms, clocks = timed(step_resident, K, ClockSampler(local) if rank == 0 else None)
    by_rank = dict(per_rank)
    host_ms = {"wait_for_batch_ms_per_step": round(1e3 * host_t["wait_batch"] / max(1, host_t["n"]), 3),
               "issue_ms_per_step": round(1e3 * host_t["issue"] / max(1, host_t["n"]), 3)}
    if world > 1:
        hm = torch.tensor([host_ms["wait_for_batch_ms_per_step"], host_ms["issue_ms_per_step"]], device=dev)
        dist.all_reduce(hm, op=dist.ReduceOp.MAX)
        host_ms["max_over_ranks"] = [round(float(hm[0]), 3), round(float(hm[1]), 3)]
    launches = _lib.launch_count - launches0
    if native:
        assert trainer.native_steps == n0 + K, "every timed step must have taken the native path"
        # The step overlaps its independent branches on up to four streams, so a kernel's CUDA-event time inside the timed region would
        # include the SMs it shares with co-scheduled kernels.  The dominant kernel is therefore timed live in a SECOND region: the same
        # K steps issued on one stream (umpr_step_streams(1)), event pairs around that entry point only.
        NS.serialize(True)
        feed["it"] = stream_of(devb, K)
        NS.profile_begin(top)
        ms_serial, _ = timed(step_resident, K)
        kt = NS.profile_end()[top]
        NS.serialize(False)
        # the 4 rotating batches differ slightly in their token counts: algorithmic work of the timed launches = K x their mean
        wk = [trainer.native.work_table(b, NS.plans_of(b, dev)).get(top, [0.0, 0.0]) for b in devb]
        kt["flops"], kt["bytes"] = K * sum(x[0] for x in wk) / NB, K * sum(x[1] for x in wk) / NB
    else:
        kt = _lib.stop_timing()[top]
    # host time to ISSUE one step with an empty launch queue (the figure above includes the back-pressure of a full queue: the host
    # runs ahead of the GPU until the driver's queue is full and is then paced by the GPU)
    feed["it"] = stream_of(devb, 3)
    time.sleep(0.05)                                       # let the prefetch workers finish the three plans
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        trainer.train_step(next(feed["it"]))
    host_ms["issue_ms_per_step_empty_queue"] = round(1e3 * (time.perf_counter() - t0) / 3, 3)
    torch.cuda.synchronize()
    value = world * B * K / (ms / 1e3)

    # ---- end to end through the public API with host buffers
    feed["it"] = stream_of(host, 2)
    for i in range(2):
        step_e2e(i)
    feed["it"] = stream_of(host, K)
    n_before = len(sink) + reader.inflight
    ms_e2e, _ = timed(lambda i: (step_e2e(i), sink.extend(reader.drain()) if i == K - 1 else None), K)
    assert len(sink) + reader.inflight - n_before == K and all(v == v for v in sink), "every step's loss must have been read back"
    e2e_val = world * B * K / (ms_e2e / 1e3)
    h2d = tensor_bytes([host[0][j] for j in (0, 1, 2, 6, 7)])
    # plus the int32 pack plans built from the host lengths (3 or 5 GRU plans share 3 buffers)
    from umpr_b200.plan import PackPlan
    for j in ((3, 4) if syn.WORKLOADS[args.workload]["review_net_only"] else (3, 4, 5)):
        pl = PackPlan(host[0][j], host[0][j - 3].shape[2], "cpu")
        h2d += pl.host.numel() * 4
        if pl.R == 128 and pl.L <= 128:   # large sides also ship their valid-row tables (S-Net / co-attention / GEMM rows, convolution tiles)
            h2d += pl._snet_host()[0].size * 4 + (pl._cnet_host()[0].size * 4 if pl.L + 2 <= 128 else 0)

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload),
                   "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "tokens_per_step_per_gpu": int(sum(tokens) / NB), "trainable_params": n_params,
                   "step": "zero_grad+fwd+bwd+allreduce+adam",
                   "gradient_exchange": ("none (1 GPU)" if world == 1 else
                                         ("C-ABI NCCL communicator (umpr_comm_*), 2 buckets [head/attention/conv/C-Net | R-Net GRU], the first all-reduced under the last backward kernel"
                                          if trainer.overlap else ("C-ABI NCCL communicator (umpr_comm_*): one all-reduce of the flat bucket after the backward" if trainer.comm is not None else "torch.distributed.all_reduce"))),
                   "host_cores_per_rank": len(cores) if cores else None, "issuing_thread_cpus": issue_cpus,
                   "host_thread": host_ms, "by_rank": by_rank or None,
                   "issue": ("one native C-ABI call per step (umpr_step: R-Net / C-Net branches, the item side of the C-Net tails and S-Net on side streams) + all-reduce + umpr_adam_step"
                             if native else "autograd Functions over per-kernel C-ABI calls"),
                   "host_pipeline": "the next batch's pack plans (torch.sort + int32 plan) are built on a worker thread, like a collate worker; rebuilt every step",
                   "l2_policy": "per-step working set (GBs of activations) exceeds the 126 MB L2; 4 rotating input batches"},
        "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4), "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    # ---- rooflines (algorithmic work per SURVEY.md §8d ÷ CUDA-event time on the launching stream).  `roofline`: the dominant
    # entry point, event-timed inside the timed region; `rooflines`: the kernels north_star names, from the all-kernel-timed step.
    ncu = {}
    tp = os.path.join(ROOT, "profiles", "ncu_metrics.json")      # per-launch ncu counters copied from profiles/ (static evidence)
    if os.path.exists(tp):
        ncu = json.load(open(tp))

    def roof(name, t, in_region):
        bound = "tensor" if name in TENSOR_BOUND else "hbm"
        calls = max(1, t["calls"])
        per_launch_ms = t["ms"] / calls
        # t["flops"] / t["bytes"]: algorithmic work summed over these calls
        if bound == "tensor":
            achieved, peak, unit = t["flops"] / calls / (per_launch_ms * 1e-3) / 1e12, peaks["tf_sust"], "TFLOP/s"
        else:
            achieved, peak, unit = t["bytes"] / calls / (per_launch_ms * 1e-3) / 1e9, peaks["hbm"], "GB/s"
        m = ncu.get(name, {})
        r = {"kernel": name, "bound": bound, "achieved": round(achieved, 3), "peak": peak, "unit": unit, "frac": round(achieved / peak, 5),
             "traffic": m.get("dram_bytes_per_launch"), "peak_source": peaks["src"] + (" bf16 dense sustained" if bound == "tensor" else " copy"),
             "launch_ms": round(per_launch_ms, 4), "launches_per_step": t["calls"] // (K if in_region else 1),
             "math": MATH.get(name, "fp32 CUDA cores")}
        if "tensor_pipe_active_pct" in m:
            r["ncu_tensor_pipe_active_pct"] = m["tensor_pipe_active_pct"]        # sm__pipe_tensor_cycles_active (3 issued MMAs per algorithmic one)
        if "dram_pct" in m:
            r["ncu_dram_pct"] = m["dram_pct"]
        return r

    line["roofline"] = roof(top, kt, True)
    if native:
        line["roofline"]["share_of_step"] = round(kt["ms"] / ms_serial, 4)
        line["roofline"]["timed"] = (f"{K} steps issued on ONE stream after the timed region ({round(ms_serial / K, 4)} ms per step): the timed region itself "
                                     "overlaps the step's branches on up to four streams, where a kernel's event time includes co-scheduled kernels")
    else:
        line["roofline"]["share_of_step"] = round(kt["ms"] / ms, 4)
    line["rooflines"] = [roof(k, table_ms[k], False) for k in NAMED if k in table_ms]
    line["kernel_table_ms"] = {k: round(v["ms"], 3) for k, v in sorted(table_ms.items(), key=lambda kv: -kv[1]["ms"])[:8]}

    if world == 1 and rank == 0:
        if not args.no_batch64:
            small = [resident(syn.make_batch(args.workload, 64, seed=77 + i)) for i in range(NB)]
            feed["it"] = stream_of(small, 3)
            for i in range(3):
                step_resident(i)
            feed["it"] = stream_of(small, K)
            ms64, _ = timed(step_resident, K)
            line["batch64"] = {"value": round(64 * K / (ms64 / 1e3), 2), "unit": UNIT, "ms_per_step": round(ms64 / K, 4),
                               "note": "reference default batch_size=64 (config.py:12), inputs resident"}
        if not args.no_cpu_baseline:
            # the reference's CPU path on this box's host cores, same workload and SAME batch as the GPU arm (a bounded sample: 4 steps),
            # and at the reference's default batch_size=64 (config.py:12) next to it
            cores = host_threads()
            v, msc, kind = cpu_reference_run(args.workload, B, 4, 1)
            v64, ms64c, _ = cpu_reference_run(args.workload, 64, 10, 2)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": round(msc, 2),
                                    "sample": f"4 train steps of batch {B} of the same workload (fwd+bwd+Adam, "
                                              + ("unmodified src/model.py from baseline/_ref" if kind == "reference" else "oracle port, torch CPU GRU")
                                              + f"), os.cpu_count={os.cpu_count()}",
                                    "batch64": {"value": round(v64, 3), "ms_per_step": round(ms64c, 2), "sample": "10 train steps of batch 64"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
