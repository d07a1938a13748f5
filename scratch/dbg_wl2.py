import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpr_b200 import synthetic as syn, functional as F
from oracle import umpr_oracle as orc
DEV = "cuda:0"
table = syn.make_table(3000, seed=2)
WL, BB, seed = sys.argv[1].split(":")
bt = syn.make_batch(WL, int(BB), vocab=3000, seed=int(seed))
model = syn.build_model(WL, table, seed=1, device=DEV)
with torch.no_grad():
    model.review_net.r_net.M.mul_(0.05)
model.train()
pred, loss = model(*bt)
loss.backward()
params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
p_ref, _, ref = orc.umpr_loss_and_grads(params, bt, review_net_only=False, impl="lib")
g = dict(model.named_parameters())
for k in ("control_net.c_net.cnn.0.weight", "control_net.c_net.cnn.0.bias", "control_net.c_net.linear.0.weight", "control_net.c_net.linear.0.bias",
          "control_net.s_net.Ms", "control_net.ss_net.linear.0.weight", "visual_net.pos_v_emb", "linear_fusion.0.weight"):
    d = (g[k].grad.cpu() - ref[k])
    print(k, "max|ref| %.3e max|err| %.3e" % (float(ref[k].abs().max()), float(d.abs().max())))
d = (g["control_net.c_net.cnn.0.bias"].grad.cpu() - ref["control_net.c_net.cnn.0.bias"]).abs()
print("bias err top filters:", d.topk(5))
print("bias ref at those:", ref["control_net.c_net.cnn.0.bias"][d.topk(5).indices])
dw = (g["control_net.c_net.cnn.0.weight"].grad.cpu() - ref["control_net.c_net.cnn.0.weight"]).abs().amax(dim=(1, 2))
print("weight err per filter top:", dw.topk(5))
