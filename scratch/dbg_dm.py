import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpr_b200 import synthetic as syn, functional as F
from oracle import umpr_oracle as orc
DEV = "cuda:0"
table = syn.make_table(3000, seed=2)
WL, BB = sys.argv[1], int(sys.argv[2])
bt = syn.make_batch(WL, BB, vocab=3000, seed=5)
ref = None
for name, flags in [("all on", {}), ("snet off", dict(TENSOR_CORE_SNET=False)), ("gru off", dict(TENSOR_CORE_GRU=False)), ("conv off", dict(TENSOR_CORE_CONV=False)), ("gemm off", dict(TENSOR_CORE_GEMM=False)), ("wgrad off", dict(TENSOR_CORE_WGRAD=False)),
                    ("all off", dict(TENSOR_CORE_GRU=False, TENSOR_CORE_GEMM=False, TENSOR_CORE_COATTN=False, TENSOR_CORE_CONV=False))]:
    saved = {k: getattr(F, k) for k in flags}
    for k, v in flags.items():
        setattr(F, k, v)
    model = syn.build_model(WL, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(0.05)
    model.train()
    pred, loss = model(*bt)
    loss.backward()
    if ref is None:
        params = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
        bt64 = tuple(b.double() if b.is_floating_point() else b for b in bt)
        _, _, ref64 = orc.umpr_loss_and_grads(params, bt64, review_net_only=False, impl="lib")
        params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        _, _, ref = orc.umpr_loss_and_grads(params, bt, review_net_only=False, impl="lib")
        ref = {k: v.double() for k, v in ref.items()}
    worst = sorted(((float((p.grad.cpu().double() - ref[k]).abs().max() / ref[k].abs().max()), k) for k, p in model.named_parameters()
                    if p.grad is not None and float(ref[k].abs().max()) > 1e-8), reverse=True)[:4]
    if name == "all on":
        w64 = sorted(((float((p.grad.cpu().double() - ref64[k]).abs().max() / ref64[k].abs().max()), k) for k, p in model.named_parameters()
                      if p.grad is not None and float(ref64[k].abs().max()) > 1e-8), reverse=True)[:4]
        print("ours vs fp64 oracle:", " | ".join(f"{k.replace('review_net.', '')} {v:.2e}" for v, k in w64), flush=True)
        o64 = sorted(((float((ref[k] - ref64[k]).abs().max() / ref64[k].abs().max()), k) for k in ref64 if float(ref64[k].abs().max()) > 1e-8), reverse=True)[:4]
        print("fp32 oracle vs fp64 oracle:", " | ".join(f"{k.replace('review_net.', '')} {v:.2e}" for v, k in o64), flush=True)
        break
    print(name, " | ".join(f"{k.replace('review_net.', '')} {v:.2e}" for v, k in worst), flush=True)
    for k, v in saved.items():
        setattr(F, k, v)
