"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU/fp32.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported from where it lies; nothing is copied.  The only
harness patch replaces ``torchvision.models.vgg16`` — which the
reference constructor would try to download — by ``nn.Flatten`` so that
``photos`` carries the backbone's 1000-d output features directly
(``(B,V,Pc,1000,1,1)`` → ``model.py:216-218`` yields ``(B,V,Pc,1000)`` unchanged).
"""
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import cases  # noqa: E402

torchvision.models.vgg16 = lambda pretrained=True, num_classes=1000: torch.nn.Flatten()
from src import model as ref  # noqa: E402


def run_umpr(name, c):
    torch.manual_seed(0)
    params = cases.make_params(c["review_net_only"], c["V"], c["vocab"], c["seed"], c["m_scale"])
    batch = cases.make_batch(c)
    m = ref.UMPR(cases.CaseConfig(c), params["embedding.weight"].numpy())
    missing = m.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    taps = {}
    m.review_net.r_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("rnet", o))
    m.review_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("represent", o))
    if not c["review_net_only"]:
        m.control_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("control", o))
        m.visual_net.register_forward_hook(lambda mod, i, o: taps.__setitem__("visual", o))
    b = list(batch)
    if not c["review_net_only"]:
        b[6] = b[6].reshape(*b[6].shape, 1, 1)
    m.train()
    pred, loss = m(*b)
    loss = loss.mean()                      # main.py:34
    m.zero_grad()
    loss.backward()
    out = {"pred": pred.detach().numpy(), "loss": loss.detach().numpy()}
    for k, p in m.named_parameters():
        if p.requires_grad:
            out["grad:" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    for i, nm in enumerate(["gru_u", "gru_i", "soft_u", "soft_i", "atte_u", "atte_i"]):
        out["rnet:" + nm] = taps["rnet"][i].detach().numpy()
    out["represent"] = taps["represent"].detach().numpy()
    if not c["review_net_only"]:
        for i, nm in enumerate(["c_u", "c_i", "prefer_pos", "prefer_neg"]):
            out["control:" + nm] = taps["control"][i].detach().numpy()
        for i, nm in enumerate(["pos_match", "neg_match", "final_pos", "final_neg"]):
            out["visual:" + nm] = taps["visual"][i].detach().numpy()
    # eval-mode forward (evaluate.py:6-14)
    m.eval()
    with torch.no_grad():
        pe, le = m(*b)
    out["eval_pred"] = pe.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    stats = {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("grad:")}
    print(name, "loss", float(loss), "pred", pred.detach().numpy().round(3)[:4],
          "min/max grad-max", min(stats.values()), max(stats.values()))
    if not c["review_net_only"]:
        print("   control", [taps["control"][i].detach().abs().max().item() for i in range(4)])


def run_rnn(name, c):
    data, lens, w, cot_out, cot_hid = cases.make_rnn_case(c)
    rnn = ref.ImprovedRnn(torch.nn.GRU, input_size=cases.E, hidden_size=cases.H, batch_first=True, bidirectional=True)
    rnn.load_state_dict(w, strict=True)
    result, hidden = rnn(data, lens)
    pk = torch.nn.utils.rnn.pack_padded_sequence(data, lens.cpu(), batch_first=True, enforce_sorted=False)
    ((result * cot_out).sum() + (hidden * cot_hid).sum()).backward()
    out = {"result": result.detach().numpy(), "hidden": hidden.detach().numpy(),
           "sorted_indices": pk.sorted_indices.numpy(), "unsorted_indices": pk.unsorted_indices.numpy(),
           "batch_sizes": pk.batch_sizes.numpy()}
    for k, p in rnn.named_parameters():
        out["grad:" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "result max", float(result.abs().max()))


def run_sort():
    rs = np.random.RandomState(5)
    out = {}
    for n in (64, 1280, 5000):
        lens = torch.tensor(rs.randint(1, 21, size=n).astype(np.int64))
        s, idx = torch.sort(lens, descending=True)
        out[f"len{n}"] = lens.numpy()
        out[f"idx{n}"] = idx.numpy()
    np.savez_compressed(os.path.join(HERE, "sort_order.npz"), **out)


def run_collate():
    """The reference's own collate (dataset.py:153-182, ignore_photos: no JPEGs here) on seeded ragged samples."""
    from src import dataset as ref_ds
    for name, c in cases.COLLATE_CASES.items():
        out = ref_ds.batch_loader(cases.make_collate_case(c), ignore_photos=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: out[i].numpy() for i, k in
                                                                    enumerate(["user", "item", "ui", "u_len", "i_len", "ui_len", "photos", "labels"])})
        print(name, [tuple(t.shape) for t in out])


if __name__ == "__main__":
    torch.set_num_threads(4)
    run_collate()
    for name, c in cases.RNN_CASES.items():
        run_rnn(name, c)
    for name, c in cases.CASES.items():
        run_umpr(name, c)
    run_sort()
    print("torch", torch.__version__)
