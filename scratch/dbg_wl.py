import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpr_b200 import synthetic as syn, functional as F
from oracle import umpr_oracle as orc
DEV = "cuda:0"
table = syn.make_table(3000, seed=2)
for spec in sys.argv[1:]:
    WL, BB, seed = spec.split(":")
    bt = syn.make_batch(WL, int(BB), vocab=3000, seed=int(seed))
    model = syn.build_model(WL, table, seed=1, device=DEV)
    with torch.no_grad():
        model.review_net.r_net.M.mul_(0.05)
    model.train()
    pred, loss = model(*bt)
    loss.backward()
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    p_ref, _, ref = orc.umpr_loss_and_grads(params, bt, review_net_only=False, impl="lib")
    worst = sorted(((float((p.grad.cpu() - ref[k]).abs().max() / ref[k].abs().max()), k) for k, p in model.named_parameters()
                    if p.grad is not None and float(ref[k].abs().max()) > 1e-7), reverse=True)[:3]
    print(spec, "pred", float((pred.cpu() - p_ref).abs().max()), " | ".join(f"{k.replace('review_net.', '').replace('control_net.', '')} {v:.2e}" for v, k in worst), flush=True)
