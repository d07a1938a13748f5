"""R-Net pre-training module (reference pretrain/pretrain_rnet.py:144-169): the only other caller of ``RNet`` - one sentence per
sample (S = 1), co-attention over its L words, a 256 -> 1 sigmoid head and BCE loss.  Same kernels as the main path; the head is one
fused kernel each way (csrc/pretrain.cu).  Dataset building (ABAE labels, gensim) is out of scope."""
import torch
from torch import nn

from . import functional as F
from .model import PackedReviews, RNet, _on


class PretrainRNet(nn.Module):
    """pretrain_rnet.py:144-169: ``forward(u (B,L) ids, u_length (B,), i, i_length, target (B,)) -> (result (B,), loss)``."""

    def __init__(self, word_emb, gru_hidden):
        super().__init__()
        self.embedding = nn.Embedding.from_pretrained(torch.as_tensor(word_emb, dtype=torch.float32).clone())
        self.r_net = RNet(self.embedding.embedding_dim, gru_hidden)
        self.linear = nn.Sequential(nn.Linear(gru_hidden * 4, 1), nn.Sigmoid())
        self.loss_fn = nn.BCELoss()          # parameter-free; kept so the attribute tree matches the reference

    def forward(self, u, u_length, i, i_length, target):
        table = self.embedding.weight
        device = table.device
        if device.type != "cuda":
            raise RuntimeError("umpr_b200: this path runs on CUDA only (no CPU fallback); move the model to a GPU")
        with _on(device):
            u, i, target = [d.to(device) for d in (u, i, target)]                       # pretrain_rnet.py:157
            pu = PackedReviews(u_length.view(-1, 1), ids=u.view(u.shape[0], 1, u.shape[1]), table=table)    # :158-161, embedding gather fused
            pi = PackedReviews(i_length.view(-1, 1), ids=i.view(i.shape[0], 1, i.shape[1]), table=table)
            _, _, _, _, att_u, att_i = self.r_net(pu, pi, None, None)                    # :164
            lin = self.linear[0]
            return F.bce_head(att_u, att_i, lin.weight, lin.bias, target.to(torch.float32))   # :165-168

    def save_r_net(self, save_path):
        torch.save(self.r_net, save_path)
