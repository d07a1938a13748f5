"""Parity of the CUDA path (through the C-ABI) against the reference's golden outputs and the CPU oracle.
Bar: bit-exact packing order / zero pattern; max|a-b|/max|b| <= 1e-4 per fp32 tensor."""
import numpy as np
import pytest
import torch

import cases
from conftest import assert_close, load_golden, rel_max
from oracle import umpr_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


def _check_grad(k, got, ref):
    ref = torch.as_tensor(ref)
    if k == "visual_net.linear.bias":
        # the bias cancels in (pos_emb - img_emb): the analytic gradient is exactly 0 and both sides hold only fp32
        # cancellation noise of O(1e-7) -> absolute comparison
        assert float(got.abs().max()) < 1e-5 and float(ref.abs().max()) < 1e-5, k
    elif float(ref.abs().max()) < 1e-7:
        assert float(got.abs().max()) < 1e-6, k
    else:
        assert_close(got, ref, TOL, "grad " + k)


def _model(c):
    import umpr_b200
    params = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    m = umpr_b200.UMPR(cases.CaseConfig(c), params["embedding.weight"])
    r = m.load_state_dict(params, strict=True)          # reference state_dict keys load unchanged
    assert not r.missing_keys and not r.unexpected_keys
    return m.to(DEV), params


@pytest.mark.parametrize("name", list(cases.RNN_CASES))
@pytest.mark.parametrize("R", [32, 64, 128])
def test_improved_rnn_vs_reference(name, R, monkeypatch):
    import umpr_b200
    from umpr_b200 import plan as plan_mod
    monkeypatch.setattr(plan_mod, "choose_tile_rows", lambda n, sm: R)
    g = load_golden(name)
    data, lens, w, cot_out, cot_hid = cases.make_rnn_case(cases.RNN_CASES[name])
    rnn = umpr_b200.ImprovedRnn(torch.nn.GRU, input_size=cases.E, hidden_size=cases.H, batch_first=True, bidirectional=True)
    rnn.load_state_dict(w, strict=True)
    rnn = rnn.to(DEV)
    result, hidden = rnn(data.to(DEV), lens)            # lengths stay on the host, as in the reference
    assert_close(result, g["result"], TOL, "result")
    assert_close(hidden, g["hidden"], TOL, "hidden")
    assert np.array_equal((result.detach().cpu().numpy() != 0).any(-1), (g["result"] != 0).any(-1)), "zero pattern"
    ((result * cot_out.to(DEV)).sum() + (hidden * cot_hid.to(DEV)).sum()).backward()
    for k, p in rnn.named_parameters():
        assert_close(p.grad, g["grad:" + k], TOL, k)


def test_improved_rnn_accepts_gpu_lengths_and_total_length():
    import umpr_b200
    torch.manual_seed(0)
    rnn = umpr_b200.ImprovedRnn(torch.nn.GRU, input_size=50, hidden_size=64, batch_first=True, bidirectional=True).to(DEV)
    data = torch.randn(37, 11, 50)
    lens = torch.randint(1, 6, (37,))                    # longest sequence (5) < padded length (11)
    out, hid = rnn(data.to(DEV), lens.to(DEV))
    w = orc.gru_weights({k: v.detach().cpu() for k, v in rnn.state_dict().items()}, "module")
    ref, hid_ref = orc.improved_rnn(data, lens, w)
    assert out.shape == (37, 11, 128)
    assert_close(out, ref, TOL, "out")
    assert_close(hid, hid_ref, TOL, "hidden")
    assert float(out.detach()[:, 5:].abs().max()) == 0.0


@pytest.mark.parametrize("name", list(cases.CASES))
def test_umpr_vs_reference(name):
    c = cases.CASES[name]
    g = load_golden(name)
    m, _ = _model(c)
    batch = cases.make_batch(c)
    m.train()
    taps = {}
    m.review_net.r_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("rnet", o))
    m.review_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("represent", o))
    if not c["review_net_only"]:
        m.control_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("control", o))
        m.visual_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("visual", o))
    pred, loss = m(*batch)
    loss = loss.mean()                                    # main.py:34
    m.zero_grad()
    loss.backward()
    for i, nm in enumerate(["gru_u", "gru_i", "soft_u", "soft_i", "atte_u", "atte_i"]):
        assert_close(taps["rnet"][i], g["rnet:" + nm], TOL, nm)
    assert np.array_equal((taps["rnet"][0].detach().cpu().numpy() != 0), (g["rnet:gru_u"] != 0)), "gru_u zero pattern"
    assert_close(taps["represent"], g["represent"], TOL, "represent")
    if not c["review_net_only"]:
        for i, nm in enumerate(["c_u", "c_i", "prefer_pos", "prefer_neg"]):
            assert_close(taps["control"][i], g["control:" + nm], TOL, nm)
        for i, nm in enumerate(["pos_match", "neg_match", "final_pos", "final_neg"]):
            assert_close(taps["visual"][i], g["visual:" + nm], TOL, nm)
    assert_close(pred, g["pred"], TOL, "pred")
    assert_close(loss, g["loss"], TOL, "loss")
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        ref = g["grad:" + k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        _check_grad(k, got, ref)
    m.eval()
    with torch.no_grad():
        pe, _ = m(*batch)
    assert_close(pe, g["eval_pred"], TOL, "eval_pred")


@pytest.mark.parametrize("workload,B", [("music_small_r", 16), ("music_full", 12), ("yelp_full", 9)])
def test_umpr_vs_oracle_amazon_shape(workload, B):
    """configs[0..2] shapes (S=20, L=20, GloVe-50d; V=1 or 4) at a batch the oracle finishes in seconds."""
    from umpr_b200 import synthetic as syn
    table = syn.make_table(5000, seed=2)
    batch = syn.make_batch(workload, B, vocab=5000, seed=7)
    m = syn.build_model(workload, table, seed=1, device=DEV)
    with torch.no_grad():
        m.review_net.r_net.M.mul_(0.05)                  # un-saturate tanh so grad(M) is measurable
    m.train()
    pred, loss = m(*batch)
    loss.backward()
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    rno = syn.WORKLOADS[workload]["review_net_only"]
    p_ref, l_ref, g_ref = orc.umpr_loss_and_grads(params, batch, review_net_only=rno, impl="explicit")
    assert_close(pred, p_ref, TOL, "pred")
    assert_close(loss, l_ref, TOL, "loss")
    for k, p in m.named_parameters():
        if p.requires_grad:
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            _check_grad(k, got, g_ref[k])


def test_module_level_api_rnet_snet_cnet():
    """RNet / SNet / CNet called the way pretrain_rnet.py:165 and model.py:162,182 call them (dense embeddings in)."""
    import umpr_b200
    torch.manual_seed(3)
    B, S, L = 7, 3, 9
    emb_u, emb_i = torch.randn(B, S, L, 50) * 0.5, torch.randn(B, S, L, 50) * 0.5
    lu, li = torch.randint(1, L + 1, (B, S)), torch.randint(1, L + 1, (B, S))
    rnet = umpr_b200.RNet(50, 64).to(DEV)
    with torch.no_grad():
        rnet.M.mul_(0.05)
    out = rnet(emb_u.to(DEV), emb_i.to(DEV), lu, li)
    p = {"review_net.r_net." + k: v.detach().cpu() for k, v in rnet.state_dict().items()}
    ref = orc.r_net(emb_u, emb_i, lu, li, p)
    for a, b, nm in zip(out, ref, ["gru_u", "gru_i", "soft_u", "soft_i", "atte_u", "atte_i"]):
        assert_close(a, b, TOL, nm)
    snet = umpr_b200.SNet(64, 128).to(DEV)
    sa, se = snet(out[0], out[2], L)
    sa_ref, se_ref = orc.s_net(ref[0], ref[2], L, snet.Ms.detach().cpu(), snet.Ws.detach().cpu())
    assert_close(sa, sa_ref, TOL, "self_atte")
    assert_close(se, se_ref, TOL, "sentiment")
    cnet = umpr_b200.CNet(50, 64, 120, 3, 4, 0.35).to(DEV)
    g, vp, fr = cnet(emb_u.to(DEV), lu)
    pc = {"control_net.c_net." + k: v.detach().cpu() for k, v in cnet.state_dict().items()}
    g_ref, vp_ref, fr_ref = orc.c_net(emb_u, lu, pc, 0.35)
    assert_close(g, g_ref, TOL, "cnet gru")
    assert_close(vp, vp_ref, TOL, "view_p")
    assert_close(fr, fr_ref, TOL, "final_repr")


def test_shard_lengths_use_global_total_length():
    """A data-parallel shard whose own longest sentence is shorter than L must still return (.., L, ..) (readme.md:154-160)."""
    from umpr_b200 import synthetic as syn
    table = syn.make_table(3000, seed=4)
    user, item, ui, ul, il, uil, photos, labels = syn.make_batch("music_small_r", 8, vocab=3000, seed=5)
    ul = ul.clamp(max=7)
    il = il.clamp(max=7)
    user = user * (torch.arange(20)[None, None] < ul[..., None])
    item = item * (torch.arange(20)[None, None] < il[..., None])
    m = syn.build_model("music_small_r", table, seed=2, device=DEV)
    m.eval()
    with torch.no_grad():
        out = m.review_net.r_net(umpr_b200_packed(m, user, ul), umpr_b200_packed(m, item, il), ul, il)
        pred, loss = m(user, item, ui, ul, il, uil, photos, labels)
    assert out[0].shape == (8, 400, 128)
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    p_ref, l_ref = orc.umpr_forward(params, (user, item, ui, ul, il, uil, photos, labels), review_net_only=True)
    assert_close(pred, p_ref, TOL, "pred")


def umpr_b200_packed(m, ids, lens):
    from umpr_b200 import PackedReviews
    return PackedReviews(lens, ids=ids.to(DEV), table=m.embedding.weight)


def test_error_behaviour():
    import umpr_b200
    rnn = umpr_b200.ImprovedRnn(torch.nn.GRU, input_size=50, hidden_size=64, batch_first=True, bidirectional=True).to(DEV)
    with pytest.raises(RuntimeError, match="greater than 0"):
        rnn(torch.randn(3, 4, 50, device=DEV), torch.tensor([2, 0, 1]))
    from umpr_b200 import synthetic as syn
    m = syn.build_model("music_small_r", syn.make_table(100), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(*syn.make_batch("music_small_r", 2, vocab=100))


def test_two_shard_step_matches_chunked_oracle():
    """SURVEY.md §8e parity check on one GPU: each DataParallel chunk alone == oracle on that chunk (own sort, global L),
    and the bucket sum / k == mean over chunks of the oracle gradients; then the fused Adam equals torch.optim.Adam."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import FlatTrainer, shard_batch
    table = syn.make_table(3000, seed=9)
    batch = syn.make_batch("music_small_r", 11, vocab=3000, seed=10)
    m = syn.build_model("music_small_r", table, seed=4, device=DEV)
    with torch.no_grad():
        m.review_net.r_net.M.mul_(0.05)
    params = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    tr = FlatTrainer(m, lr=1e-2, weight_decay=1e-3)
    tr.zero_grad()
    ref = {}
    k_shards = 2
    for r in range(k_shards):
        shard = shard_batch(batch, r, k_shards)
        assert shard[0].shape[0] == (6 if r == 0 else 5)
        pred, loss = m(*shard)
        loss.backward()                                             # accumulates into the flat bucket, like the all-reduce sum
        p_ref, l_ref, g_ref = orc.umpr_loss_and_grads(params, shard, review_net_only=True)
        assert_close(pred, p_ref, TOL, f"pred shard {r}")
        assert_close(loss, l_ref, TOL, f"loss shard {r}")
        for k, g in g_ref.items():
            ref[k] = ref.get(k, 0) + g / k_shards
    for k, p in m.named_parameters():
        if p.requires_grad:
            _check_grad(k, p.grad / k_shards, ref[k])
    # one optimizer step on the averaged gradient vs torch.optim.Adam with the reference's parameter groups (main.py:22-25)
    # (Adam divides by |g|: feed both sides the SAME averaged gradient so the kernel, not the conditioning, is what is tested)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.requires_grad:
                p.grad.copy_(ref[k].to(DEV) * k_shards)
    tr.world = k_shards
    tr.optimizer_step()
    tp = {k: v.clone().requires_grad_(True) for k, v in params.items() if k != "embedding.weight"}
    opt = torch.optim.Adam([{"params": [v for k, v in tp.items() if "bias" not in k]},
                            {"params": [v for k, v in tp.items() if "bias" in k], "weight_decay": 0.0}], 1e-2, weight_decay=1e-3)
    for k, v in tp.items():
        v.grad = ref[k].clone()
    opt.step()
    for k, p in m.named_parameters():
        if p.requires_grad:
            assert_close(p.detach(), tp[k].detach(), 2e-5, "adam " + k)


@pytest.mark.parametrize("workload,batch,m_scale,seed", [("music_full", 128, 1.0, 5), ("music_full", 128, 0.05, 5),
                                                         ("yelp_full", 112, 0.05, 5), ("yelp_full", 120, 0.05, 7), ("music_full", 112, 0.05, 5)])
def test_full_model_on_the_tensor_core_path_vs_oracle(workload, batch, m_scale, seed):
    """A batch large enough (B*S >= 2048 sentences) that every tensor-core kernel is on the path - fused tcgen05 GRU forward and
    backward with all sides in one launch, tcgen05 co-attention and S-Net, TN reduction GEMM, sweep conv backward - against the CPU
    oracle: prediction, loss and every parameter gradient within the 1e-4 bar, at the reference's own initialisation (m_scale 1:
    tanh saturated, arg-max gradients ~0) and with M scaled down so that the co-attention gradients are exercised.

    A batch of this size holds ~1e5 co-attention maxima and ~1e6 max-pool maxima; a handful are tied to within the last bits, so
    two fp32 implementations route those gradients to different (equally maximal) positions and disagree by up to 1e-2 on grad(M)
    and the conv weights - whatever their precision.  The oracle therefore back-propagates through the positions OUR kernels
    chose (oracle.routed) and the test asserts separately that every chosen position is a maximum to within 2e-5 of the value
    range."""
    from umpr_b200 import functional as F
    from umpr_b200 import synthetic as syn
    table = syn.make_table(3000, seed=2)
    batch_t = syn.make_batch(workload, batch, vocab=3000, seed=seed)
    assert batch_t[3].numel() >= 2048
    model = syn.build_model(workload, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(m_scale)
    model.train()
    F.ROUTING_LOG = []
    try:
        pred, loss = model(*batch_t)
        log = F.ROUTING_LOG
    finally:
        F.ROUTING_LOG = None
    loss.backward()
    picks = {"coattn": [(a[0].cpu(), a[1].cpu()) for k, a in log if k == "coattn"], "cnet": [a.cpu() for k, a in log if k == "cnet"]}
    assert len(picks["coattn"]) == 1 and len(picks["cnet"]) == 3
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    rno = syn.WORKLOADS[workload]["review_net_only"]
    with orc.routed(picks) as r:
        p_ref, l_ref, g_ref = orc.umpr_loss_and_grads(params, batch_t, review_net_only=rno, impl="lib")
    assert r.margin["coattn"] <= 2e-5 and r.margin["cnet"] <= 2e-5, r.margin     # our winners ARE maxima (to the element error)
    assert_close(pred, p_ref, TOL, "prediction")
    assert_close(loss, l_ref, TOL, "loss")
    for k, p in model.named_parameters():
        if p.requires_grad:
            _check_grad(k, p.grad if p.grad is not None else torch.zeros_like(p), g_ref[k])


def test_abi_allreduce_two_gpus():
    """C-ABI NCCL helpers (umpr_comm_unique_id / umpr_comm_init / umpr_allreduce, SURVEY.md §8b,e) on two ranks: same sums as
    torch.distributed, and a FlatTrainer step through them ends with identical parameters.  Needs two GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(here, "dist_abi_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ABI_ALLREDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("B,L", [(40, 20), (2100, 12)])
def test_pretrain_rnet_vs_oracle(B, L):
    """the R-Net pre-training module (pretrain_rnet.py:144-169; one sentence per sample, sigmoid head, BCE) through
    the same kernels - small batch on the CUDA-core path, 2100 sentences on the fused tensor-core path - against the oracle."""
    from umpr_b200.pretrain import PretrainRNet
    from umpr_b200 import synthetic as syn
    torch.manual_seed(B)
    table = syn.make_table(3000, seed=2)
    m = PretrainRNet(table, 64).to(DEV)
    with torch.no_grad():
        m.r_net.M.mul_(0.05)
    lens = [torch.randint(1, L + 1, (B,)) for _ in range(2)]
    ids = [torch.where(torch.arange(L)[None, :] < ln[:, None], torch.randint(3, 3000, (B, L)), torch.zeros(B, L, dtype=torch.int64)) for ln in lens]
    target = torch.randint(0, 2, (B,)).float()
    from umpr_b200 import functional as F
    m.train()
    F.ROUTING_LOG = []
    try:
        result, loss = m(ids[0], lens[0], ids[1], lens[1], target)
        log = F.ROUTING_LOG
    finally:
        F.ROUTING_LOG = None
    loss.mean().backward()                                                   # pretrain_rnet.py:190-192
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    p = {k: v.clone().requires_grad_(k != "embedding.weight") for k, v in params.items()}
    # thousands of co-attention maxima: the oracle back-propagates through the positions our kernels chose (see the large-batch test)
    with orc.routed({"coattn": [(a[0].cpu(), a[1].cpu()) for k, a in log if k == "coattn"]}) as r:
        r_ref, l_ref = orc.pretrain_rnet_forward(p, ids[0], lens[0], ids[1], lens[1], target, impl="lib")
    assert r.margin["coattn"] <= 2e-5, r.margin
    keys = [k for k in p if k != "embedding.weight"]
    g_ref = dict(zip(keys, torch.autograd.grad(l_ref, [p[k] for k in keys], allow_unused=True)))
    assert_close(result, r_ref.detach(), TOL, "result")
    assert_close(loss, l_ref.detach(), TOL, "loss")
    for k, prm in m.named_parameters():
        if prm.requires_grad:
            ref = g_ref[k] if g_ref[k] is not None else torch.zeros_like(p[k])
            got = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            if k == "linear.0.bias":
                # a single number: mean(p - t) over the batch, ~1e-2 of the mean magnitude of its terms - both fp32 sums (ours and the
                # oracle's) are only good to ~1e-6 of THAT scale, so the bar is applied to it, not to the cancelled result
                scale = float((r_ref.detach() - target).abs().mean())
                assert float((got.cpu() - ref).abs().max()) <= TOL * scale, (k, float(got), float(ref), scale)
            else:
                _check_grad(k, got, ref)


def test_standalone_ssnet_trains():
    """SSNet (model.py:129-143) called on its own: sigmoid(Linear(128 -> 1)), forward and all three gradients against torch."""
    import umpr_b200
    torch.manual_seed(8)
    net = umpr_b200.SSNet(128).to(DEV)
    x = (torch.randn(6, 5, 128, device=DEV) * 0.7).requires_grad_(True)
    gy = torch.randn(6, 5, 1, device=DEV)
    y = net(x)
    assert y.shape == (6, 5, 1)
    (y * gy).sum().backward()
    w, b = net.linear[0].weight.detach().clone().requires_grad_(True), net.linear[0].bias.detach().clone().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.sigmoid(xr @ w.t() + b)
    (yr * gy).sum().backward()
    assert_close(y, yr, 1e-6, "ssnet forward")
    assert_close(x.grad, xr.grad, 1e-5, "ssnet dx")
    assert_close(net.linear[0].weight.grad, w.grad, 1e-5, "ssnet dw")
    assert_close(net.linear[0].bias.grad, b.grad, 1e-5, "ssnet db")


def test_unwritten_gradient_rows_are_never_read():
    """The kernels leave the gradient rows of positions beyond a sentence's length unwritten (the packed GRU backward never reads
    them, model.py:18).  Filled with NaN instead, every parameter gradient must come out the same."""
    from umpr_b200 import functional as F
    from umpr_b200 import synthetic as syn
    table = syn.make_table(3000, seed=2)
    batch = syn.make_batch("music_full", 24, vocab=3000, seed=11)          # 480 sentences per side: tensor-core path with plans
    res = []
    for poison in (False, True):
        model = syn.build_model("music_full", table, seed=1, device=DEV)
        with torch.no_grad():
            model.review_net.r_net.M.mul_(0.05)
        F.POISON_UNWRITTEN = poison
        try:
            pred, loss = model(*batch)
            loss.backward()
        finally:
            F.POISON_UNWRITTEN = False
        res.append({k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad and p.grad is not None})
    for k in res[0]:
        assert torch.isfinite(res[1][k]).all(), k
        if float(res[0][k].abs().max()) > 1e-7:
            assert_close(res[1][k], res[0][k], 2e-6, "poisoned " + k)       # float atomics: last-bit differences only


def test_snet_with_a_foreign_word_soft_keeps_its_input_gradient():
    """functional.GradSink hands S-Net's input gradient to the co-attention backward ONLY when word_soft is that node's soft-max.
    With any other word_soft (here: a fresh tensor) S-Net must return its dx itself - the GRU output's gradient is then the sum of
    both consumers' contributions, exactly what plain autograd gives."""
    import umpr_b200
    torch.manual_seed(3)
    B, S, L = 40, 8, 9                                                   # 320 sentences: tensor-core kernels with plans
    emb_u, emb_i = torch.randn(B, S, L, 50) * 0.5, torch.randn(B, S, L, 50) * 0.5
    lu, li = torch.randint(1, L + 1, (B, S)), torch.randint(1, L + 1, (B, S))
    rnet = umpr_b200.RNet(50, 64).to(DEV)
    snet = umpr_b200.SNet(64, 128).to(DEV)
    with torch.no_grad():
        rnet.M.mul_(0.05)
    grads = []
    for foreign in (False, True):
        rnet.zero_grad(); snet.zero_grad()
        out = rnet(emb_u.to(DEV), emb_i.to(DEV), lu, li)
        ws = torch.softmax(torch.randn(B, S * L, device=DEV, generator=torch.Generator(DEV).manual_seed(1)), -1).requires_grad_(True) if foreign else out[2]
        sa, se = snet(out[0], ws, L)
        ((se ** 2).sum() + out[4].sum()).backward()
        grads.append({k: p.grad.clone() for k, p in rnet.named_parameters()})
    p = {"review_net.r_net." + k: v.detach().cpu() for k, v in rnet.state_dict().items()}
    # oracle for the foreign case: the same graph in CPU autograd
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = orc.r_net(emb_u, emb_i, lu, li, pr, impl="lib")
    ws_c = torch.softmax(torch.randn(B, S * L, device=DEV, generator=torch.Generator(DEV).manual_seed(1)), -1).cpu()
    _, se_ref = orc.s_net(ref[0], ws_c, L, snet.Ms.detach().cpu(), snet.Ws.detach().cpu())
    ((se_ref ** 2).sum() + ref[4].sum()).backward()
    for k, g in grads[1].items():
        r = pr["review_net.r_net." + k].grad
        if float(r.abs().max()) > 1e-7:
            assert_close(g, r, TOL, "foreign word_soft: grad " + k)


def test_model_on_a_non_current_device():
    """The reference lets the model live on any ``config.device``; every forward makes that device current for its kernels."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from umpr_b200 import synthetic as syn
    table = syn.make_table(3000, seed=2)
    batch = syn.make_batch("music_full", 8, vocab=3000, seed=3)
    torch.cuda.set_device(0)
    out = []
    for dev in ("cuda:0", "cuda:1"):
        m = syn.build_model("music_full", table, seed=1, device=dev)
        pred, loss = m(*batch)
        loss.backward()
        out.append((pred.detach().cpu(), {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}))
    assert torch.cuda.current_device() == 0
    assert_close(out[1][0], out[0][0], 1e-6, "prediction on cuda:1")
    for k, g in out[0][1].items():
        if float(g.abs().max()) > 1e-7:
            assert_close(out[1][1][k], g, 1e-5, "grad on cuda:1 " + k)


@pytest.mark.parametrize("name", list(cases.COLLATE_CASES))
def test_device_collate_matches_reference_batch_loader(name):
    """umpr_b200.data.collate (ragged lists shipped flat, padded on the device by umpr_collate_ids) against the reference's
    batch_loader fixtures: bit-exact ids (int64, on the device) and lengths (on the host, >= 1)."""
    from umpr_b200.data import collate
    g = load_golden(name)
    out = collate(cases.make_collate_case(cases.COLLATE_CASES[name]), DEV, ignore_photos=True)
    for i, k in enumerate(["user", "item", "ui"]):
        assert out[i].is_cuda and out[i].dtype == torch.int64
        assert np.array_equal(out[i].cpu().numpy(), g[k]), k
    for i, k in zip((3, 4, 5), ["u_len", "i_len", "ui_len"]):
        assert not out[i].is_cuda and np.array_equal(out[i].numpy(), g[k]), k
    assert out[6].numel() == 0 and np.array_equal(out[7].numpy(), g["labels"])


def test_feature_cache_feeds_the_visual_tail(tmp_path):
    """photo ids -> rows of the on-disk VGG16 feature table -> umpr_feature_gather -> VisualNet tail.  The gathered
    tensor equals oracle.photo_features (missing photos take the zero image's features), and the model's forward + backward on a
    collated batch equals the one on explicitly supplied features."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.data import FeatureStore, collate
    rs = np.random.RandomState(7)
    n_photo, V, Pc, B = 300, 4, 2, 9
    ids = ["ph%d" % i for i in range(n_photo)]
    feats = (rs.normal(0, 0.05, size=(n_photo, 1000))).astype(np.float32)
    zero_img = (rs.normal(0, 0.05, size=1000)).astype(np.float32)
    store = FeatureStore.build(str(tmp_path / "vgg16.umprfeat"), ids, feats, missing_features=zero_img).to_device(DEV)
    pick = rs.randint(0, n_photo, size=(B, V, Pc))
    names = [[["unknown" if rs.rand() < 0.2 else "ph%d" % pick[b, v, p] for p in range(Pc)] for v in range(V)] for b in range(B)]
    rows = store.rows_of(names)
    got = store.gather(rows)
    want = orc.photo_features(torch.from_numpy(np.concatenate([feats, zero_img[None]])), torch.from_numpy(rows), store.missing_row)
    assert torch.equal(got.cpu(), want)
    # through the model: collate with the store vs. the same features passed explicitly
    c = cases.COLLATE_CASES["collate_a"]
    ragged = (cases.make_collate_case(c) + cases.make_collate_case(dict(c, seed=77)))[:B]
    samples = [(s[0], s[1], s[2], names[b], s[4]) for b, s in enumerate(ragged)]
    batch = collate(samples, DEV, feature_store=store)
    table = syn.make_table(600, seed=2)
    m = syn.build_model("yelp_full", table, seed=1, device=DEV)
    pred, loss = m(*batch)
    loss.backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    explicit = list(batch)
    explicit[6] = want.reshape(B, V, Pc, 1000)
    pred2, loss2 = m(*explicit)
    loss2.backward()
    assert torch.equal(pred, pred2) or rel_max(pred, pred2) < 1e-6
    for k, p in m.named_parameters():
        if p.grad is not None and float(p.grad.abs().max()) > 1e-7:
            assert_close(g1[k], p.grad, 1e-5, "grad " + k)
