import sys, torch
sys.path.insert(0, '/root/repo')
from umpr_b200 import synthetic as syn, functional as F
from umpr_b200._lib import call, ptr
import umpr_b200
DEV='cuda:0'
table = syn.make_table(5000, seed=2)
batch = syn.make_batch('yelp_full', 9, vocab=5000, seed=7)
m = syn.build_model('yelp_full', table, seed=1, device=DEV)
cn = m.control_net.c_net
from umpr_b200.model import PackedReviews
for name, ids, lens in (('ui', batch[2], batch[5]), ('u', batch[0], batch[3]), ('i', batch[1], batch[4])):
    pk = PackedReviews(lens, ids=ids.to(DEV), table=m.embedding.weight)
    with torch.no_grad():
        g, _ = cn.gru.run(pk, want_hidden=False)
        x = g.view(pk.B, pk.S*pk.L, 128).contiguous()
        N, L, KC = pk.B*pk.S, pk.L, 120
        outs = {}
        for tcflag in (False, True):
            F.TENSOR_CORE_CONV = tcflag
            cw, cb = cn.cnn[0].weight.detach(), cn.cnn[0].bias.detach()
            cfeat = torch.empty(N, KC, device=DEV); cidx = torch.empty(N, KC, dtype=torch.int32, device=DEV)
            if tcflag:
                cap = max(4096, N*KC//8)
                scratch = torch.zeros((197632 + 16*cap)//4, device=DEV)
                call("umpr_cnet_conv_fwd_tc", ptr(x), ptr(cw), ptr(cb), N, L, KC, 3, ptr(scratch), cap, ptr(cfeat), ptr(cidx), 148)
                torch.cuda.synchronize()
                cnt = scratch.view(torch.int32)[(196608+512)//4].item()
                print(name, 'records', cnt, 'of', N*KC)
            else:
                wt = torch.empty(3*128*128, device=DEV)
                call("umpr_cnet_prep", ptr(cw), KC, 3, ptr(wt))
                call("umpr_cnet_conv_fwd", ptr(x), ptr(wt), ptr(cb), N, L, KC, ptr(cfeat), ptr(cidx), 148)
            outs[tcflag] = (cfeat.clone(), cidx.clone())
        d = (outs[True][1] != outs[False][1])
        print(name, 'cidx mismatches', int(d.sum()), 'max cfeat diff', float((outs[True][0]-outs[False][0]).abs().max()))
        # reference conv in fp64
        xx = x.view(N, L, 128).double().transpose(1,2)
        y = torch.nn.functional.conv1d(xx, cw.double(), cb.double(), padding=1)
        for n, k in d.nonzero()[:10].tolist():
            a, b_ = outs[True][1][n,k].item(), outs[False][1][n,k].item()
            print('   n', n, 'kf', k, 'tc', a, 'fp32', b_, 'y64', [round(v,7) for v in y[n,k].tolist()][:6], ' y[a],y[b]=', y[n,k,max(a,0)].item(), y[n,k,max(b_,0)].item(), 'len', lens.reshape(-1)[:0].tolist())
