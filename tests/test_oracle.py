"""The oracle against the reference's own outputs (tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

import cases
from conftest import assert_close, load_golden
from oracle import umpr_oracle as orc

TOL = 2e-6   # two fp32 CPU evaluation orders of the same arithmetic


@pytest.mark.parametrize("name", list(cases.RNN_CASES))
@pytest.mark.parametrize("impl", ["explicit", "lib"])
def test_improved_rnn_matches_reference(name, impl):
    g = load_golden(name)
    data, lens, w, cot_out, cot_hid = cases.make_rnn_case(cases.RNN_CASES[name])
    ws = [w_.clone().requires_grad_(True) for w_ in orc.gru_weights(w, "module")]
    sorted_len, sorted_idx, unsorted = orc.sort_plan(lens)
    assert np.array_equal(sorted_idx.numpy(), g["sorted_indices"])          # bit-exact packing order
    assert np.array_equal(unsorted.numpy(), g["unsorted_indices"])
    result, hidden = orc.improved_rnn(data, lens, ws, impl=impl)
    assert_close(result, g["result"], TOL, "result")
    assert_close(hidden, g["hidden"], TOL, "hidden")
    # zero pattern: row n is non-zero exactly for t < len[unsorted[n]]
    eff = lens[unsorted]
    mask = torch.arange(data.shape[1])[None, :] < eff[:, None]
    assert torch.equal((result.detach().abs().sum(-1) > 0), mask)
    assert np.array_equal((np.abs(g["result"]).sum(-1) > 0), mask.numpy())
    ((result * cot_out).sum() + (hidden * cot_hid).sum()).backward()
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    names = names + [n + "_reverse" for n in names]
    for n_, w_ in zip(names, ws):
        assert_close(w_.grad, g["grad:module." + n_], 2e-5, n_)


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("impl", ["explicit", "lib"])
def test_umpr_matches_reference(name, impl):
    c = cases.CASES[name]
    g = load_golden(name)
    params = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    batch = cases.make_batch(c)
    pred, loss, grads = orc.umpr_loss_and_grads(params, batch, review_net_only=c["review_net_only"], impl=impl)
    assert_close(pred, g["pred"], TOL, "pred")
    assert_close(loss, g["loss"], TOL, "loss")
    for k, v in grads.items():
        ref = g["grad:" + k]
        if np.abs(ref).max() < 1e-7:
            assert float(v.abs().max()) < 1e-6, k
        else:
            assert_close(v, ref, 5e-5, "grad " + k)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_intermediates_match_reference(name):
    c = cases.CASES[name]
    g = load_golden(name)
    p = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    user, item, ui, ul, il, uil, photos, labels = cases.make_batch(c)
    t = p["embedding.weight"]
    out = orc.r_net(t[user], t[item], ul, il, p)
    for nm, v in zip(["gru_u", "gru_i", "soft_u", "soft_i", "atte_u", "atte_i"], out):
        assert_close(v, g["rnet:" + nm], TOL, nm)
    assert_close(orc.review_net(t[user], t[item], ul, il, p), g["represent"], TOL, "represent")
    if not c["review_net_only"]:
        cu, ci, pp, pn = orc.control_net(t[user], t[item], t[ui], ul, il, uil, p, 0.35)
        for nm, v in zip(["c_u", "c_i", "prefer_pos", "prefer_neg"], (cu, ci, pp, pn)):
            assert_close(v, g["control:" + nm], 5e-6, nm)
        vis = orc.visual_net_tail(photos, cu, ci, p)
        for nm, v in zip(["pos_match", "neg_match", "final_pos", "final_neg"], vis):
            assert_close(v, g["visual:" + nm], 5e-6, nm)


def test_sort_order_is_the_reference_call():
    g = load_golden("sort_order")
    for n in (64, 1280, 5000):
        _, idx, _ = orc.sort_plan(torch.tensor(g[f"len{n}"]))
        assert np.array_equal(idx.numpy(), g[f"idx{n}"])


def test_length_zero_is_an_error():
    with pytest.raises(RuntimeError):
        orc.sort_plan(torch.tensor([3, 0, 2]))


def test_adam_restatement_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(37)
    g = torch.randn(37)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3, weight_decay=1e-3)
    m, v, p = torch.zeros(37), torch.zeros(37), p0.clone()
    for step in range(1, 4):
        p_ref.grad = g.clone()
        opt.step()
        orc.adam_step(p, g, m, v, step, 1e-3, 1e-3)
    assert_close(p, p_ref.detach(), 1e-6, "adam")
