"""Parity at the configurations that are BENCHMARKED and claimed (BASELINE.json configs[1..4]), against the CPU oracle.

* the bench configuration itself - music_full at per-GPU batch 1024 (360 GRU tiles over 148 slot queues: 2-3 tiles per queue, so the
  backward kernel's weight-gradient accumulation in tensor memory spans tiles) - plus yelp_full (V=4) and csj_long (full sentence pools)
  at batch 256;
* the fused tensor-core GRU forward + backward with artificially deep tile queues (2 and 5 CTAs: up to 9 tiles per queue), against the
  CUDA-core kernels AND against ``oracle.improved_rnn``;
* the evaluation forward (evaluate.py:6-14) on length-skewed batches at padded sentence lengths 32 / 64 / 128, including one sample whose
  valid positions exceed 512 (the limit of one tensor-core co-attention pass).

Arg-max routing: the oracle back-propagates through the positions our kernels chose (``oracle.routed``).  That is only
sound if those positions ARE maxima and if almost all of them are the oracle's own arg-max, so the tests also
  (1) bound the margin of every chosen position (<= 2e-5 of the value range),
  (2) count the positions that differ from the oracle's own arg-max, and bound the fraction whose value is strictly lower (near-ties
      decided differently; exact ties - saturated tanh = 1.0, padded zeros - carry no gradient),
  (3) assert that with exactly those entries masked out of the gradient the oracle under OUR routing and under ITS OWN routing agree,
      i.e. the routing differs nowhere else,
  (4) hold prediction, loss and every routing-INDEPENDENT parameter gradient to the 1e-4 bar against the plain, unrouted oracle.
"""
import json
import os

import pytest
import torch

from conftest import ROOT, assert_close, rel_max
from oracle import umpr_oracle as orc
from test_gpu_parity import _check_grad

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


def _routing_dependent(key: str) -> bool:
    """Parameters upstream of a max (co-attention row/column maxima, C-Net max-pool): their gradient depends on which of several
    near-tied positions wins.  Everything else only sees the (continuous) maximum VALUES."""
    return key.endswith("r_net.M") or ".gru.module." in key or ".cnn.0." in key


def _record(name, payload):
    """Keep the routing statistics of a run next to the other GPU artefacts (gpurun_out/ is merged back by gpurun)."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "parity_routing_stats.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = payload
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass


def _run_model(workload, batch, m_scale, seed, vocab=30000, native=False):
    """One forward + backward of the CUDA path: through the autograd Functions (``model(*batch)``; ``loss.backward()``), or - ``native`` -
    through the one-call native step (csrc/step.cu) exactly as ``FlatTrainer.train_step`` and bench.py run it."""
    from umpr_b200 import functional as F
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import FlatTrainer
    table = syn.make_table(vocab, seed=2)
    batch_t = syn.make_batch(workload, batch, vocab=vocab, seed=seed)
    model = syn.build_model(workload, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(m_scale)
    model.train()
    log = []
    if native:
        tr = FlatTrainer(model)
        tr.zero_grad()
        plans = tr.native.plans_of(batch_t, torch.device(DEV))
        assert tr.native.supported(batch_t, plans)
        pred, loss = tr.native.run(batch_t, True, plans, routing_log=log)
    else:
        F.ROUTING_LOG = log
        try:
            pred, loss = model(*batch_t)
        finally:
            F.ROUTING_LOG = None
        loss.backward()
    picks = {"coattn": [(a[0].cpu(), a[1].cpu()) for k, a in log if k == "coattn"], "cnet": [a.cpu() for k, a in log if k == "cnet"]}
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu() for k, p in model.named_parameters() if p.requires_grad}
    return model, batch_t, picks, params, pred.detach().cpu(), loss.detach().cpu(), grads


@pytest.mark.parametrize("workload,batch,m_scale,masked,native", [("music_full", 1024, 1.0, False, True), ("yelp_full", 256, 0.05, True, True),
                                                                  ("csj_long", 256, 0.05, True, True), ("music_small_r", 512, 0.05, False, True),
                                                                  ("music_full", 512, 0.05, False, False)])
def test_benchmarked_configurations_vs_oracle(workload, batch, m_scale, masked, native):
    from umpr_b200 import synthetic as syn
    model, batch_t, picks, params, pred, loss, grads = _run_model(workload, batch, m_scale, seed=5, native=native)
    rno = syn.WORKLOADS[workload]["review_net_only"]
    kw = dict(review_net_only=rno, impl="lib")
    with orc.routed(picks) as r:
        p_ref, l_ref, g_ref = orc.umpr_loss_and_grads(params, batch_t, **kw)
    stats = {"margin": r.margin, "stats": r.stats}
    # (1) our winners are maxima, (2) and almost always the oracle's own
    assert r.margin["coattn"] <= 2e-5 and r.margin["cnet"] <= 2e-5, r.margin
    for kind, st in r.stats.items():
        if st["total"]:
            stats[kind + "_strict_fraction"] = st["strict"] / st["total"]
            assert st["strict"] <= 2e-4 * st["total"], (kind, st)
    # parity under the common routing: prediction, loss, EVERY parameter gradient
    assert_close(pred, p_ref, TOL, "prediction")
    assert_close(loss, l_ref, TOL, "loss")
    for k, g in grads.items():
        _check_grad(k, g, g_ref[k])
    # (4) the plain oracle, no routing: forward values and every gradient that does not hinge on the winner of a near-tie
    p_own, l_own, g_own = orc.umpr_loss_and_grads(params, batch_t, **kw)
    assert_close(pred, p_own, TOL, "prediction (unrouted)")
    assert_close(loss, l_own, TOL, "loss (unrouted)")
    unrouted = {}
    for k, g in grads.items():
        if _routing_dependent(k):
            unrouted[k] = rel_max(g, g_own[k]) if float(g_own[k].abs().max()) > 1e-7 else 0.0
        else:
            _check_grad(k, g, g_own[k])
    stats["unrouted_rel_err_of_routing_dependent_grads"] = unrouted
    if masked:
        # (3) masked comparison: our routing vs the oracle's own routing, no gradient through the entries where they differ
        res = {}
        for mode in ("routed_masked", "own_masked"):
            pk = {k: list(v) for k, v in picks.items()}
            with orc.routed(pk, mode=mode):
                res[mode] = orc.umpr_loss_and_grads(params, batch_t, **kw)[2]
        worst = 0.0
        for k in grads:
            if float(res["own_masked"][k].abs().max()) > 1e-7:
                e = rel_max(res["routed_masked"][k], res["own_masked"][k])
                worst = max(worst, e)
                assert e <= TOL, f"masked routing comparison, grad {k}: {e:.3e}"
        stats["masked_worst_rel_err"] = worst
    _record(f"{workload}_b{batch}_m{m_scale}_{'native' if native else 'autograd'}", stats)


def _weights(E, seed):
    torch.manual_seed(seed)
    gru = torch.nn.GRU(E, 64, batch_first=True, bidirectional=True)
    return [p.detach().clone() for p in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, gru.weight_ih_l0_reverse,
                                         gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse)]


@pytest.mark.parametrize("ctas", [2, 5])
def test_fused_gru_autograd_with_deep_tile_queues(ctas):
    """Three ImprovedRnn calls in one fused launch, scheduled onto 2 / 5 CTAs per direction: every slot queue holds several tiles, so
    the forward's slot hand-over between tiles and the backward's weight-gradient accumulation in tensor memory ACROSS tiles are what
    runs.  Outputs, h_n and all eight weight gradients against the fp32 CUDA-core kernels and against oracle.improved_rnn."""
    from umpr_b200 import functional as F
    from umpr_b200.plan import PackPlan
    torch.manual_seed(ctas)
    E = 50
    sides = [(3000, 20), (1500, 20), (700, 12)]
    w0 = _weights(E, 11)
    lens = [torch.randint(1, L + 1, (N,)) for N, L in sides]
    for ln in lens:
        ln[torch.rand(ln.numel()) < 0.3] = 1
    data = [torch.randn(N, L, E) * 0.5 for N, L in sides]
    gy = [torch.randn(N, L, 128) for N, L in sides]
    gh = [torch.randn(2, N, 64) for N, L in sides]
    plans = [PackPlan(ln, L, DEV, tile_rows=128) for ln, (N, L) in zip(lens, sides)]
    xps = [F.gather_pack(p, dense=d.to(DEV))[0] for p, d in zip(plans, data)]
    xqs = [F.gather_pack_tc(p, dense=d.to(DEV))[0] for p, d in zip(plans, data)]
    n_tiles = sum(p.n_tiles for p in plans)
    assert n_tiles >= 3 * 2 * ctas, "queues must hold at least three tiles"
    res = []
    for flag in (False, True):
        F.TENSOR_CORE_GRU, F.GRU_SCHED_CTAS = flag, ctas
        try:
            w = [t.to(DEV).requires_grad_(True) for t in w0]
            outs = F.gru_forward_multi(plans, xps, xqs, E, w, want_hidden=True)
            sum((o * g.to(DEV)).sum() + (h * q.to(DEV)).sum() for (o, h), g, q in zip(outs, gy, gh)).backward()
            res.append(([(o.detach().cpu(), h.detach().cpu()) for o, h in outs], [t.grad.cpu() for t in w]))
        finally:
            F.TENSOR_CORE_GRU, F.GRU_SCHED_CTAS = True, None
    # the CPU oracle (model.py:12-21 restated), same cotangents
    wc = [t.clone().requires_grad_(True) for t in w0]
    total = 0
    ref_out = []
    for d, ln, g, q in zip(data, lens, gy, gh):
        o, h = orc.improved_rnn(d, ln, wc, impl="lib")
        ref_out.append((o.detach(), h.detach()))
        total = total + (o * g).sum() + (h * q).sum()
    total.backward()
    for i in range(len(sides)):
        for j, nm in enumerate(("out", "hn")):
            assert torch.equal(res[1][0][i][j] == 0, ref_out[i][j] == 0) or nm == "hn", f"side {i}: zero pattern of {nm}"
            assert_close(res[1][0][i][j], ref_out[i][j], TOL, f"side {i} {nm} vs oracle")
            assert_close(res[1][0][i][j], res[0][0][i][j], 2e-5, f"side {i} {nm} vs CUDA-core path")
    for k, (a, b, c) in enumerate(zip(res[1][1], res[0][1], wc)):
        assert_close(a, c.grad, TOL, f"weight gradient {k} vs oracle")
        assert_close(a, b, 1e-4, f"weight gradient {k} vs CUDA-core path")      # two fp32-class implementations: each is within its own error of the oracle


@pytest.mark.parametrize("workload,B,L", [("music_full", 24, 32), ("music_full", 16, 64), ("music_small_r", 12, 128), ("music_full", 8, 126)])
def test_eval_forward_on_skewed_lengths_vs_oracle(workload, B, L):
    """BASELINE configs[4] shapes: evaluate.py:6-14 forward (eval, no_grad) with 90 % of the sentences 1-4 tokens long and 10 % at the
    padded length; sample 0 has every sentence at full length (20*L valid positions: beyond one 512-position co-attention pass)."""
    import numpy as np
    from umpr_b200 import synthetic as syn
    from umpr_b200.eval import evaluate_mse
    vocab = 3000
    table = syn.make_table(vocab, seed=2)
    batch = list(syn.make_batch(workload, B, vocab=vocab, seed=L, L=L, skew=True))
    rs = np.random.RandomState(L)
    for j in (0, 1):                                                     # user and item side of sample 0: all sentences full
        batch[3 + j][0] = L
        batch[j][0] = torch.from_numpy(rs.randint(3, vocab, size=tuple(batch[j][0].shape)))
    rno = syn.WORKLOADS[workload]["review_net_only"]
    model = syn.build_model(workload, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(0.05)
    model.eval()
    with torch.no_grad():
        pred, loss = model(*batch)
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    p_ref, l_ref = orc.umpr_forward(params, tuple(batch), review_net_only=rno, impl="lib")
    assert_close(pred, p_ref, TOL, "prediction")
    assert_close(loss, l_ref, TOL, "loss")
    # the package-level evaluation loop (evaluate.py:6-14): sum of squared errors over the loader / sample count
    mse = evaluate_mse(model, [tuple(batch), tuple(batch)])
    want = float(((p_ref - batch[7]) ** 2).sum()) / B
    assert abs(mse - want) <= TOL * max(1.0, want), (mse, want)


@pytest.mark.parametrize("workload,B", [("music_full", 64), ("music_small_r", 40), ("yelp_full", 33)])
def test_native_step_equals_the_autograd_path(workload, B):
    """csrc/step.cu issues the same kernels as functional.py: prediction, loss and every parameter gradient agree to float-atomics noise
    (reference default batch 64 included: 1280 sentences per side, 320 user->item sentences - all on the tensor-core path)."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import FlatTrainer
    table = syn.make_table(3000, seed=2)
    batch = syn.make_batch(workload, B, vocab=3000, seed=9)
    res = []
    for native in (False, True):
        model = syn.build_model(workload, table, seed=1, device=DEV)
        with torch.no_grad():
            model.review_net.r_net.M.mul_(0.05)
        tr = FlatTrainer(model, native=native)
        tr.zero_grad()
        if native:
            plans = tr.native.plans_of(batch, torch.device(DEV))
            assert tr.native.supported(batch, plans)
            pred, loss = tr.native.run(batch, True, plans)
        else:
            assert tr.native is None
            pred, loss = model(*batch)
            loss.backward()
        res.append((pred.detach().clone(), loss.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}))
    assert_close(res[1][0], res[0][0], 1e-6, "prediction")
    assert_close(res[1][1], res[0][1], 1e-6, "loss")
    for k, g in res[0][2].items():
        if float(g.abs().max()) > 1e-7:
            assert_close(res[1][2][k], g, 5e-6, "grad " + k)
    # and the evaluation forward (no_grad -> native) equals the autograd-path forward
    model.eval()
    with torch.no_grad():
        pe, le = model(*batch)
    from umpr_b200 import model as M
    M.NATIVE_EVAL = False
    try:
        with torch.no_grad():
            pa, la = model(*batch)
    finally:
        M.NATIVE_EVAL = True
    assert_close(pe, pa, 1e-6, "eval prediction")
    assert_close(le, la, 1e-6, "eval loss")


def test_flat_trainer_steps_through_the_native_path():
    """FlatTrainer.train_step on the standard model: the native one-call path is what runs, and a few steps of it move the parameters
    exactly as the autograd path does."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import FlatTrainer, PlanPrefetcher
    table = syn.make_table(3000, seed=2)
    batches = [syn.make_batch("music_full", 64, vocab=3000, seed=20 + i) for i in range(3)]
    flats, grads = [], []
    for native in (True, False):
        model = syn.build_model("music_full", table, seed=1, device=DEV)
        with torch.no_grad():
            model.review_net.r_net.M.mul_(0.05)
        tr = FlatTrainer(model, lr=1e-6, native=native)              # the reference's learning rate (config.py:13)
        for b in PlanPrefetcher(iter(batches), torch.device(DEV)):
            tr.train_step(b)
        assert tr.native_steps == (3 if native else 0)
        flats.append(tr.flat.clone())
        grads.append(tr.grad.clone())
    # (Adam normalises every gradient component to a step of ~lr, noise included: parameters are compared at the size of one step,
    # the last step's gradient bucket - which depends on the two earlier updates - to float-atomics noise)
    assert float((flats[0] - flats[1]).abs().max()) <= 3 * 1e-6 * 2.01
    assert float((grads[0] - grads[1]).abs().max() / grads[1].abs().max()) < 1e-5


@pytest.mark.parametrize("workload,B,L", [("music_small_r", 8, 64), ("music_full", 8, 48)])
def test_training_with_more_than_512_valid_positions_vs_oracle(workload, B, L):
    """Samples whose valid positions exceed one 512-position co-attention pass (here 20 full sentences of L tokens: 1280 / 960 valid
    positions, 10 / 8 column tiles): the tensor-core affinity kernel walks all tile pairs, candidates of a column are carried across
    the row tiles in global memory.  Forward, loss and every gradient against the oracle, routing handed over and checked."""
    import numpy as np
    from umpr_b200 import functional as F
    from umpr_b200 import synthetic as syn
    vocab = 3000
    table = syn.make_table(vocab, seed=2)
    batch = list(syn.make_batch(workload, B, vocab=vocab, seed=L, L=L))
    rs = np.random.RandomState(L)
    for j in (0, 1):                                                     # half of the samples: every sentence at full length
        batch[3 + j][: B // 2] = L
        batch[j][: B // 2] = torch.from_numpy(rs.randint(3, vocab, size=tuple(batch[j][: B // 2].shape)))
    batch = tuple(batch)
    model = syn.build_model(workload, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(0.05)
    model.train()
    F.ROUTING_LOG = []
    try:
        pred, loss = model(*batch)
        log = F.ROUTING_LOG
    finally:
        F.ROUTING_LOG = None
    loss.backward()
    picks = {"coattn": [(a[0].cpu(), a[1].cpu()) for k, a in log if k == "coattn"], "cnet": [a.cpu() for k, a in log if k == "cnet"]}
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    rno = syn.WORKLOADS[workload]["review_net_only"]
    with orc.routed(picks) as r:
        p_ref, l_ref, g_ref = orc.umpr_loss_and_grads(params, batch, review_net_only=rno, impl="lib")
    assert r.margin["coattn"] <= 2e-5 and r.margin["cnet"] <= 2e-5, r.margin
    assert_close(pred, p_ref, TOL, "prediction")
    assert_close(loss, l_ref, TOL, "loss")
    for k, p in model.named_parameters():
        if p.requires_grad:
            _check_grad(k, p.grad if p.grad is not None else torch.zeros_like(p), g_ref[k])


@pytest.mark.parametrize("workload,B", [("music_full", 512), ("music_small_r", 256), ("music_full", 48)])
def test_native_step_on_side_streams_equals_the_single_stream_step(workload, B):
    """``umpr_step`` issues its independent branches on up to three streams of the library beside the caller's (C-Net beside R-Net, the item
    side of the C-Net tails, S-Net beside the co-attention; the fourth one only when the batch is large).  ``umpr_step_streams(1)`` puts
    everything on the caller's stream: same prediction, loss and gradient bucket (to the noise of float atomics)."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.step import NativeStep
    from umpr_b200.train import FlatTrainer
    table = syn.make_table(5000, seed=2)
    batch = syn.make_batch(workload, B, vocab=5000, seed=31)
    res = []
    try:
        for serialize in (True, False, False):           # twice on the side streams: the second call reuses streams and events
            NativeStep.serialize(serialize)
            model = syn.build_model(workload, table, seed=1, device=DEV)
            with torch.no_grad():
                model.review_net.r_net.M.mul_(0.05)
            tr = FlatTrainer(model)
            tr.zero_grad()
            plans = tr.native.plans_of(batch, torch.device(DEV))
            assert tr.native.supported(batch, plans)
            pred, loss = tr.native.run(batch, True, plans)
            torch.cuda.synchronize()
            res.append((pred.clone(), loss.clone(), tr.grad.clone()))
    finally:
        NativeStep.serialize(False)
    for pred, loss, grad in res[1:]:
        assert_close(pred, res[0][0], 1e-6, "prediction")
        assert_close(loss, res[0][1], 1e-6, "loss")
        assert float((grad - res[0][2]).abs().max() / res[0][2].abs().max()) < 1e-5
