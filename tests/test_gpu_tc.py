"""tcgen05 kernels against fp64 references (the 3xBF16 split must hold the 1e-4 bar with margin)."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 64, 128), (300, 128, 128), (1000, 192, 256), (257, 100, 52), (4096, 128, 128)])
def test_tc_gemm_nt(M, N, K):
    from umpr_b200._lib import call, ptr
    torch.manual_seed(M + N + K)
    lda = (K + 3) // 4 * 4
    ldc = (N + 3) // 4 * 4
    A = torch.randn(M, lda, device=DEV)
    B = torch.randn(N, lda, device=DEV)
    C = torch.zeros(M, ldc, device=DEV)
    call("umpr_tc_gemm_nt", ptr(A), lda, ptr(B), lda, ptr(C), ldc, M, N, K, 0, None, 0, 0)
    ref = A[:, :K].double() @ B[:, :K].double().t()
    assert_close(C[:, :N], ref, 2e-5, "tc_gemm_nt")
    # accumulate + bias + tanh epilogue
    bias = torch.randn(N, device=DEV)
    C2 = torch.randn(M, ldc, device=DEV) * 0.1
    C0 = C2.clone()
    A2, B2 = A * 0.05, B * 0.5
    call("umpr_tc_gemm_nt", ptr(A2), lda, ptr(B2), lda, ptr(C2), ldc, M, N, K, 1, ptr(bias), 1, 0)
    ref2 = torch.tanh(C0[:, :N].double() + A2[:, :K].double() @ B2[:, :K].double().t() + bias.double())
    assert_close(C2[:, :N], ref2, 2e-5, "tc_gemm_nt epilogue")


@pytest.mark.parametrize("M,N,K", [(512, 128, 128), (300, 100, 52), (640, 256, 192)])
def test_tc_gemm_b_stored_kn(M, N, K):
    from umpr_b200._lib import call, ptr
    torch.manual_seed(M)
    lda, ldb, ldc = (K + 3) // 4 * 4, (N + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.randn(M, lda, device=DEV)
    B = torch.randn(K, ldb, device=DEV)
    C = torch.zeros(M, ldc, device=DEV)
    call("umpr_tc_gemm_nt", ptr(A), lda, ptr(B), ldb, ptr(C), ldc, M, N, K, 0, None, 0, 1)
    assert_close(C[:, :N], A[:, :K].double() @ B[:, :N].double(), 2e-5, "tc_gemm b_kn")


@pytest.mark.parametrize("M,N,K,b_kn,ctas", [(5000, 128, 128, 0, 148), (5000, 128, 128, 1, 7), (1300, 64, 128, 0, 3), (777, 128, 64, 1, 148),
                                              (40000, 128, 128, 1, 148), (130, 100, 50, 0, 1)])
def test_tc_gemm_weight_stationary(M, N, K, b_kn, ctas):
    from umpr_b200._lib import call, ptr
    torch.manual_seed(M + ctas)
    lda, ldc = (K + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.randn(M, lda, device=DEV)
    B = torch.randn(K, N, device=DEV) if b_kn else torch.randn(N, K, device=DEV)
    bias = torch.randn(N, device=DEV)
    C = torch.randn(M, ldc, device=DEV) * 0.1
    C0 = C.clone()
    call("umpr_tc_gemm_ws", ptr(A), lda, ptr(B), B.shape[1], ptr(C), ldc, M, N, K, 1, ptr(bias), 0, b_kn, None, 0, 0, ctas)
    Bm = B.double() if b_kn else B.double().t()
    assert_close(C[:, :N], C0[:, :N].double() + A[:, :K].double() @ Bm + bias.double(), 2e-5, "tc_gemm_ws")


@pytest.mark.parametrize("R", [32, 64, 128])
def test_inproj_tc_matches_cuda_core_path(R):
    from umpr_b200 import functional as F
    from umpr_b200._lib import call, ptr, ptr_array
    from umpr_b200.plan import PackPlan
    torch.manual_seed(R)
    N, L, E = 300, 12, 50
    lens = torch.randint(1, L + 1, (N,))
    data = torch.randn(N, L, E, device=DEV) * 0.5
    plan = PackPlan(lens, L, DEV, tile_rows=R)
    xp, _ = F.gather_pack(plan, dense=data)
    gru = torch.nn.GRU(E, 64, batch_first=True, bidirectional=True).to(DEV)
    w = [p.detach().contiguous() for p in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, gru.weight_ih_l0_reverse,
                                           gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse)]
    n = plan.n_slabs * 2 * R * 192
    G0 = torch.empty(n, device=DEV)
    G1 = torch.empty(n, device=DEV)
    call("umpr_gru_inproj", ptr(xp), ptr_array(w), plan.n_slabs, R, E, ptr(G0))
    call("umpr_gru_inproj_tc", ptr(xp), ptr_array(w), plan.n_slabs, R, E, ptr(G1), 148)
    assert_close(G1, G0, 2e-5, "inproj tc vs fp32")


def _gru_weights(E, seed):
    torch.manual_seed(seed)
    gru = torch.nn.GRU(E, 64, batch_first=True, bidirectional=True).to(DEV)
    return [p.detach().contiguous() for p in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, gru.weight_ih_l0_reverse,
                                              gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse)]


@pytest.mark.parametrize("sides,E,ctas", [([(300, 12)], 50, 74), ([(1000, 20)], 50, 2), ([(2500, 20), (2500, 20)], 50, 74),
                                          ([(640, 7), (1500, 20), (129, 20)], 50, 5), ([(20480, 20)], 50, 74), ([(700, 33)], 17, 74)])
def test_fused_gru_forward_matches_cuda_core_path(sides, E, ctas):
    """umpr_gru_fwd_tc (input projection + recurrence on tcgen05, several sides per launch) against the fp32 CUDA-core
    kernels: same output rows / zero pattern (bit-exact zeros), values and h_n within the 3xBF16 bar; in training mode the h_t operand
    images it streams out (all the backward kernel needs besides the token images) hold h_t = the output rows."""
    import ctypes as C
    from umpr_b200 import _lib, functional as F
    from umpr_b200._lib import call, ptr, ptr_array
    from umpr_b200.plan import PackPlan, build_schedule
    torch.manual_seed(len(sides) * 100 + E)
    w = _gru_weights(E, 7)
    plans, xps, refs = [], [], []
    for N, L in sides:
        lens = torch.randint(1, L + 1, (N,))
        lens[torch.rand(N) < 0.3] = 1                      # many length-1 padding sentences, as the collate produces
        data = torch.randn(N, L, E, device=DEV) * 0.5
        plan = PackPlan(lens, L, DEV, tile_rows=128)
        xp, _ = F.gather_pack(plan, dense=data)
        xq, _ = F.gather_pack_tc(plan, dense=data)
        G = torch.empty(plan.n_slabs * 2 * 128 * 192, device=DEV)
        call("umpr_gru_inproj", ptr(xp), ptr_array(w), plan.n_slabs, 128, E, ptr(G))
        out = torch.full((N, L, 128), 7.0, device=DEV)
        hn = torch.full((2, N, 64), 7.0, device=DEV)
        sv = torch.zeros(plan.n_slabs * 2 * 128 * 256, device=DEV)
        call("umpr_gru_recurrence_fwd", ptr(G), ptr_array(w), ptr(plan.buf), plan.n_tiles, plan.n_slabs, 128, N, L, ptr(out), ptr(hn), ptr(sv))
        plans.append(plan); xps.append(xq); refs.append((out, hn, sv))
    n = len(sides)
    segs = (_lib.GruSeg * n)()
    got = []
    for i, (plan, xp) in enumerate(zip(plans, xps)):
        N, L = sides[i]
        out = torch.full((N, L, 128), -3.0, device=DEV)
        hn = torch.full((2, N, 64), -3.0, device=DEV)
        hq = torch.zeros(plan.n_slabs * 2 * 2 * 128 * 128, dtype=torch.uint8, device=DEV)
        segs[i] = _lib.GruSeg(ptr(xp), ptr(plan.buf), ptr(out), ptr(hn), ptr(hq), plan.n_tiles, plan.n_slabs, N, L)
        got.append((out, hn, hq))
    sched, nq = build_schedule([p.tile_len for p in plans], ctas)
    sched = torch.from_numpy(sched).to(DEV)
    call("umpr_gru_fwd_tc", C.addressof(segs), n, ptr_array(w), E, ptr(sched), nq)
    torch.cuda.synchronize()
    for i, ((o0, h0, s0), (o1, h1, hq)) in enumerate(zip(refs, got)):
        assert torch.equal(o0 == 0, o1 == 0), f"side {i}: zero pattern differs"
        assert_close(o1, o0, 2e-5, f"side {i} out")
        assert_close(h1, h0, 2e-5, f"side {i} hn")
        # hq[slab = tile_off + t][dir][hi|lo][row][64 bf16, SWIZZLE_128B] holds h_t of the tile's jobs: decode job 0 of tile 0 and compare
        # with its output row (forward direction: columns 0..63)
        plan = plans[i]
        raw = hq.view(plan.n_slabs, 2, 2, 128, 128)                       # bytes: (slab, direction, hi|lo, row, 128 B)
        bf = raw[:, :, :, 0].contiguous().view(torch.bfloat16).float()    # row 0: its 16-byte chunks are not permuted (c ^ (0 & 7) = c)
        h_img = bf[:, :, 0] + bf[:, :, 1]
        r0 = int(plan.host[plan.n_tiles * 128 + 0])                       # row_of[job 0]
        L0 = int(plan.host[2 * plan.n_tiles * 128 + 0])                   # its length = the steps of tile 0
        if L0 > 1:
            # the image written at step t is h_t; the last step's image is never needed and not written
            assert_close(h_img[:L0 - 1, 0], o1[r0, :L0 - 1, :64], 2e-5, f"side {i} hq image (forward direction)")


def test_fused_gru_autograd_matches_cuda_core_path():
    from umpr_b200 import functional as F
    from umpr_b200.plan import PackPlan
    torch.manual_seed(5)
    E, N, L = 50, 3000, 20
    w0 = _gru_weights(E, 11)
    lens = torch.randint(1, L + 1, (N,))
    data = torch.randn(N, L, E, device=DEV) * 0.5
    plan = PackPlan(lens, L, DEV, tile_rows=128)
    xp, _ = F.gather_pack(plan, dense=data)
    xq, _ = F.gather_pack_tc(plan, dense=data)
    gy = torch.randn(N, L, 128, device=DEV)
    res = []
    for flag in (False, True):
        F.TENSOR_CORE_GRU = flag
        w = [t.clone().requires_grad_(True) for t in w0]
        out, _ = F.gru_forward(plan, xp, E, w, want_hidden=False, xq=xq)
        (out * gy).sum().backward()
        res.append((out.detach(), [t.grad.clone() for t in w]))
    F.TENSOR_CORE_GRU = True
    assert_close(res[1][0], res[0][0], 2e-5, "out")
    for a, b in zip(res[1][1], res[0][1]):
        assert_close(a, b, 5e-5, "grad")


@pytest.mark.parametrize("M,N,K,ctas", [(128, 128, 5000, 148), (128, 128, 64, 3), (64, 100, 777, 7), (128, 128, 409600, 148), (4, 128, 1000, 1)])
def test_tc_gemm_tn_reduction(M, N, K, ctas):
    """C += A^T · B with the reduction index slowest in both operands (dM = gi^T · dgiM, model.py:50 backward)."""
    from umpr_b200._lib import call, ptr
    torch.manual_seed(M + N + K)
    lda, ldb, ldc = 128, 128, (N + 3) // 4 * 4
    A = torch.randn(K, lda, device=DEV) * 0.3
    B = torch.randn(K, ldb, device=DEV) * 0.3
    C = torch.randn(M, ldc, device=DEV)
    C0 = C.clone()
    call("umpr_tc_gemm_tn", ptr(A), lda, ptr(B), ldb, ptr(C), ldc, M, N, K, ctas)
    ref = C0[:, :N].double() + A[:, :M].double().t() @ B[:, :N].double()
    assert_close(C[:, :N], ref, 2e-5, "tc_gemm_tn")
    if ldc > N:
        assert torch.equal(C[:, N:], C0[:, N:])


@pytest.mark.parametrize("B,S,L,KC", [(128, 16, 20, 120), (9, 3, 11, 100)])
def test_cnet_tail_tensor_core_conv_matches_cuda_core_conv(B, S, L, KC):
    """C-Net tail (conv + ReLU + max-pool + view head, model.py:105-125) with the tcgen05 implicit-GEMM convolution against the
    fp32 CUDA-core convolution: same pooled features and the same gradients (the arg-max routing must not differ)."""
    from umpr_b200 import functional as F
    torch.manual_seed(B + L)
    x0 = torch.randn(B, S * L, 128, device=DEV) * 0.5
    p0 = [torch.randn(KC, 128, 3, device=DEV) * 0.05, torch.randn(KC, device=DEV) * 0.1,
          torch.randn(4, KC, device=DEV) * 0.1, torch.randn(4, device=DEV) * 0.1]
    gv, gf = torch.randn(B, S, 4, device=DEV), torch.randn(B, 4, device=DEV)
    res = []
    for flag in (False, True):
        F.TENSOR_CORE_CONV = flag
        x = x0.clone().requires_grad_(True)
        p = [t.clone().requires_grad_(True) for t in p0]
        view_p, final = F.c_net_tail(x, S, L, *p, 0.35)
        ((view_p * gv).sum() + (final * gf).sum()).backward()
        res.append([view_p.detach(), final.detach(), x.grad] + [t.grad for t in p])
    F.TENSOR_CORE_CONV = True
    for a, b, nm in zip(res[1], res[0], ["view_p", "final", "dx", "d conv_w", "d conv_b", "d lin_w", "d lin_b"]):
        assert_close(a, b, 2e-5, nm)


@pytest.mark.parametrize("B,S,L,short", [(1024, 20, 20, False), (128, 20, 20, False), (300, 5, 20, True), (40, 3, 128, False), (64, 4, 100, True), (7, 2, 1, False)])
def test_snet_tensor_core_matches_cuda_core_path(B, S, L, short):
    """S-Net (model.py:71-81) on tcgen05 over the valid positions only against the fp32 kernel over all positions: same outputs,
    same parameter gradients, same input gradient on every position below a sentence's length (the others are never read)."""
    from umpr_b200 import functional as F
    from umpr_b200.plan import PackPlan
    torch.manual_seed(B + S + L)
    N = B * S
    lens = torch.randint(1, (min(4, L) if short else L) + 1, (N,))
    if short:
        lens[::7] = L
    plan = PackPlan(lens, L, DEV, tile_rows=128)
    eff = plan.row_lengths().to(DEV)                                  # length of the sequence each OUTPUT row holds
    mask = (torch.arange(L, device=DEV)[None, :] < eff[:, None]).view(B, S * L, 1)
    x0 = torch.randn(B, S * L, 128, device=DEV) * 0.5 * mask
    ws0 = torch.softmax(torch.randn(B, S * L, device=DEV), dim=1).view(B, S, L)
    p0 = [torch.randn(64, 128, device=DEV) * 0.1, torch.randn(1, 64, device=DEV)]
    g_sa, g_se = torch.randn(B, S, 128, device=DEV), torch.randn(B, 128, device=DEV)
    res = []
    for pl in (None, plan):
        x, ws = x0.clone().requires_grad_(True), ws0.clone().requires_grad_(True)
        p = [t.clone().requires_grad_(True) for t in p0]
        sa, se = F.s_net(x, ws, L, *p, plan=pl)
        ((sa * g_sa).sum() + (se * g_se).sum()).backward()
        res.append([sa.detach(), se.detach(), torch.where(mask, x.grad, torch.zeros_like(x.grad)), ws.grad, p[0].grad, p[1].grad])
    assert plan.snet_table()[1] <= (int(lens.sum()) + 128 - L) // (129 - L) + 1
    for a, b, nm in zip(res[1], res[0], ["self_atte", "sentiment", "dx", "d word_soft", "dMs", "dWs"]):
        assert_close(a, b, 5e-5, nm)


@pytest.mark.parametrize("B,S,L,KC,V", [(112, 20, 20, 120, 4), (128, 20, 20, 120, 1), (37, 5, 20, 120, 4), (9, 3, 11, 100, 2)])
def test_cnet_tail_vs_fp64_torch(B, S, L, KC, V):
    """C-Net tail (model.py:118-125) forward and backward against the same ops written with torch in fp64 (conv1d, ReLU, max over
    positions, Linear+Sigmoid, threshold, sum of squares)."""
    from umpr_b200 import functional as F
    torch.manual_seed(B * 7 + V)
    x0 = torch.randn(B, S * L, 128, device=DEV) * 0.5
    p0 = [torch.randn(KC, 128, 3, device=DEV) * 0.05, torch.randn(KC, device=DEV) * 0.1,
          torch.randn(V, KC, device=DEV) * 0.1, torch.randn(V, device=DEV) * 0.1]
    gv, gf = torch.randn(B, S, V, device=DEV), torch.randn(B, V, device=DEV)
    x = x0.clone().requires_grad_(True)
    p = [t.clone().requires_grad_(True) for t in p0]
    view_p, final = F.c_net_tail(x, S, L, *p, 0.35)
    ((view_p * gv).sum() + (final * gf).sum()).backward()
    got = [view_p.detach(), final.detach(), x.grad] + [t.grad for t in p]
    xd = x0.double().requires_grad_(True)
    pd = [t.double().requires_grad_(True) for t in p0]
    conv = torch.relu(torch.nn.functional.conv1d(xd.view(B * S, L, 128).transpose(-1, -2), pd[0], pd[1], padding=1))
    feat = conv.max(dim=-1)[0].reshape(B, S, -1)
    vp = torch.sigmoid(feat @ pd[2].t() + pd[3])
    vp = torch.where(vp < 0.35, torch.zeros_like(vp), vp)
    fin = (vp ** 2).sum(-2)
    ((vp * gv.double()).sum() + (fin * gf.double()).sum()).backward()
    ref = [vp.detach(), fin.detach(), xd.grad] + [t.grad for t in pd]
    for a, b, nm in zip(got, ref, ["view_p", "final", "dx", "d conv_w", "d conv_b", "d lin_w", "d lin_b"]):
        assert_close(a, b.float(), 3e-5, nm)


@pytest.mark.parametrize("B,S,L,short", [(96, 20, 20, False), (33, 5, 20, True), (8, 4, 100, False), (5, 3, 7, True), (12, 10, 100, True)])
def test_coattention_over_valid_rows_matches_dense(B, S, L, short):
    """Co-attention (model.py:50-55) with the pack plans of its inputs - only the valid rows are multiplied, the padded rows' exact 0
    enters every maximum analytically - against the same kernels run densely over all P rows: same soft-max weights, pooled
    vectors and gradients (input gradients compared on the valid rows; the others are never read).  The last case has P = 1000
    padded positions but fewer than 512 valid ones per sample: with the plans it runs on the tensor cores, without on the fp32 kernels."""
    from umpr_b200 import functional as F
    from umpr_b200.plan import PackPlan
    torch.manual_seed(B + L)
    N, P = B * S, S * L
    plans, xs, masks = [], [], []
    for side in range(2):
        lens = torch.randint(1, (min(3, L) if short else L) + 1, (N,))
        if short:
            lens[side::5] = L
        pl = PackPlan(lens, L, DEV, tile_rows=128 if N >= 2048 else 32)
        m = (torch.arange(L, device=DEV)[None, :] < pl.row_lengths().to(DEV)[:, None]).view(B, P, 1)
        plans.append(pl); masks.append(m)
        xs.append(torch.tanh(torch.randn(B, P, 128, device=DEV)) * m)          # GRU outputs lie in (-1, 1)
    M0 = torch.randn(128, 128, device=DEV) * (-0.02 if short else 0.02)       # negative M: many maxima are the padded rows' 0
    g = [torch.randn(B, P, device=DEV), torch.randn(B, P, device=DEV), torch.randn(B, 128, device=DEV), torch.randn(B, 128, device=DEV)]
    res = []
    for pls in (None, tuple(plans)):
        gu, gi, M = xs[0].clone().requires_grad_(True), xs[1].clone().requires_grad_(True), M0.clone().requires_grad_(True)
        out = F.co_attention(gu, gi, M, plans=pls)
        sum((o * w).sum() for o, w in zip(out, g)).backward()
        res.append([o.detach() for o in out] + [torch.where(masks[0], gu.grad, torch.zeros_like(gu.grad)), torch.where(masks[1], gi.grad, torch.zeros_like(gi.grad)), M.grad])
    for a, b, nm in zip(res[1], res[0], ["soft_u", "soft_i", "atte_u", "atte_i", "dgu", "dgi", "dM"]):
        assert_close(a, b, 5e-5 if P > 512 else 2e-5, nm)        # P > 512: the dense side is the fp32 kernel, not the same 3xBF16 arithmetic


@pytest.mark.parametrize("B,S,L,short", [(128, 20, 20, False), (200, 5, 20, True), (10, 3, 100, False), (6, 2, 3, True)])
def test_cnet_tail_over_valid_rows_matches_dense(B, S, L, short):
    """C-Net tail with the pack plan of its input - the tensor-core convolution lays out the valid rows only, the all-zero windows
    enter the max as the bias - against the same kernels over every position: same pooled views and gradients."""
    from umpr_b200 import functional as F
    from umpr_b200.plan import PackPlan
    torch.manual_seed(B * 3 + L)
    N, KC, V = B * S, 120, 4
    lens = torch.randint(1, (min(3, L) if short else L) + 1, (N,))
    if short:
        lens[::4] = L
    plan = PackPlan(lens, L, DEV, tile_rows=128 if N >= 2048 else 32)
    mask = (torch.arange(L, device=DEV)[None, :] < plan.row_lengths().to(DEV)[:, None]).view(B, S * L, 1)
    x0 = torch.tanh(torch.randn(B, S * L, 128, device=DEV)) * mask
    p0 = [torch.randn(KC, 128, 3, device=DEV) * 0.05, torch.randn(KC, device=DEV) * 0.3,      # biases of both signs: the zero windows win often
          torch.randn(V, KC, device=DEV) * 0.1, torch.randn(V, device=DEV) * 0.1]
    gv, gf = torch.randn(B, S, V, device=DEV), torch.randn(B, V, device=DEV)
    res = []
    for pl in (None, plan):
        x = x0.clone().requires_grad_(True)
        p = [t.clone().requires_grad_(True) for t in p0]
        view_p, final = F.c_net_tail(x, S, L, *p, 0.35, plan=pl)
        ((view_p * gv).sum() + (final * gf).sum()).backward()
        res.append([view_p.detach(), final.detach(), torch.where(mask, x.grad, torch.zeros_like(x.grad))] + [t.grad for t in p])
    for a, b, nm in zip(res[1], res[0], ["view_p", "final", "dx", "d conv_w", "d conv_b", "d lin_w", "d lin_b"]):
        assert_close(a, b, 5e-5 if nm in ("d conv_w", "dx") else 1e-6, nm)      # with the plan both conv gradients run on tcgen05 (3xBF16)


@pytest.mark.parametrize("Nsent,L,accumulate,ctas", [(3000, 20, 0, 148), (3000, 20, 1, 148), (5000, 7, 1, 5), (300, 128, 0, 148)])
def test_tc_gemm_weight_stationary_over_valid_rows(Nsent, L, accumulate, ctas):
    """umpr_tc_gemm_ws with a sentence-length table: the rows below each sentence's length equal the dense product, every other
    row of C is left untouched."""
    from umpr_b200._lib import call, ptr
    from umpr_b200.plan import PackPlan
    torch.manual_seed(Nsent + L)
    lens = torch.randint(1, L + 1, (Nsent,))
    plan = PackPlan(lens, L, DEV, tile_rows=128)
    table, n_tiles = plan.snet_table()
    valid = (torch.arange(L, device=DEV)[None, :] < plan.row_lengths().to(DEV)[:, None]).reshape(-1, 1)
    M = Nsent * L
    A = torch.randn(M, 128, device=DEV) * 0.3 * valid
    W = torch.randn(128, 128, device=DEV) * 0.1
    C0 = torch.randn(M, 128, device=DEV)
    C = C0.clone()
    call("umpr_tc_gemm_ws", ptr(A), 128, ptr(W), 128, ptr(C), 128, M, 128, 128, accumulate, None, 0, 0, ptr(table), n_tiles, L, ctas)
    ref = (A.double() @ W.double().t() + (C0.double() if accumulate else 0)).float()
    assert_close(torch.where(valid, C, torch.zeros_like(C)), torch.where(valid, ref, torch.zeros_like(ref)), 2e-5, "valid rows")
    assert torch.equal(torch.where(valid, torch.zeros_like(C), C), torch.where(valid, torch.zeros_like(C), C0)), "padded rows must stay untouched"
