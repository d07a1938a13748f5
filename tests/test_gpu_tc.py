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
    call("umpr_tc_gemm_ws", ptr(A), lda, ptr(B), B.shape[1], ptr(C), ldc, M, N, K, 1, ptr(bias), 0, b_kn, ctas)
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
