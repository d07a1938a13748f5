"""Host-side length bookkeeping (umpr_b200/plan.py) against the oracle and the reference's golden permutations."""
import numpy as np
import pytest
import torch

import cases
from conftest import load_golden
from oracle import umpr_oracle as orc
from umpr_b200.plan import PackPlan, choose_tile_rows


@pytest.mark.parametrize("name", list(cases.RNN_CASES))
@pytest.mark.parametrize("R", [32, 64, 128])
def test_plan_matches_reference_packing(name, R):
    g = load_golden(name)
    data, lens, *_ = cases.make_rnn_case(cases.RNN_CASES[name])
    p = PackPlan(lens, data.shape[1], "cpu", tile_rows=R)
    assert np.array_equal(p.sorted_indices.numpy(), g["sorted_indices"])          # bit-exact packing order
    assert np.array_equal(p.unsorted_indices.numpy(), g["unsorted_indices"])
    # batch_sizes of the reference PackedSequence == number of jobs alive at each step
    bs = [(p.sorted_lengths > t).sum().item() for t in range(int(p.sorted_lengths[0]))]
    assert bs == g["batch_sizes"].tolist()
    N, Rp = p.N, p.n_tiles * R
    h = p.host.to(torch.int64)
    seq_of, row_of, len_of = h[:Rp], h[Rp:2 * Rp], h[2 * Rp:3 * Rp]
    tile_off = h[3 * Rp:3 * Rp + p.n_tiles + 1]
    slab_tile = h[3 * Rp + p.n_tiles + 1:]
    assert slab_tile.numel() == p.n_slabs == int(tile_off[-1])
    # every output row is produced exactly once, from the sequence the reference's double un-sort selects
    assert sorted(row_of[:N].tolist()) == list(range(N))
    assert torch.equal(p.unsorted_indices[row_of[:N]], seq_of[:N])                # result[n] = Y[unsorted[n]]
    assert torch.equal(len_of[:N], lens[seq_of[:N]])
    assert (len_of[N:] == 0).all() and (row_of[N:] == -1).all()
    assert (len_of[:-1] >= len_of[1:]).all()                                      # descending → tiles early-exit
    for j in range(p.n_tiles):
        assert tile_off[j + 1] - tile_off[j] == len_of[j * R]
        assert (slab_tile[tile_off[j]:tile_off[j + 1]] == j).all()
    # zero pattern of the result: row n non-zero exactly for t < len[unsorted[n]]
    eff = p.row_lengths()
    assert np.array_equal((np.abs(g["result"]).sum(-1) > 0), (torch.arange(data.shape[1])[None] < eff[:, None]).numpy())


def test_plan_rejects_bad_lengths():
    with pytest.raises(RuntimeError, match="greater than 0"):
        PackPlan(torch.tensor([3, 0, 2]), 5, "cpu", tile_rows=32)
    with pytest.raises(RuntimeError, match="exceeds"):
        PackPlan(torch.tensor([3, 9, 2]), 5, "cpu", tile_rows=32)


def test_total_length_is_the_padded_dim_not_the_shard_max():
    # readme.md:154-160: a DataParallel shard whose longest sentence is shorter than the padded tensor
    lens = torch.tensor([2, 1, 3, 1])
    p = PackPlan(lens, 20, "cpu", tile_rows=32)
    assert p.L == 20 and p.n_slabs == 3


def test_sort_is_the_reference_call():
    g = load_golden("sort_order")
    for n in (64, 1280, 5000):
        p = PackPlan(torch.tensor(g[f"len{n}"]), 20, "cpu", tile_rows=128)
        assert np.array_equal(p.sorted_indices.numpy(), g[f"idx{n}"])
        assert torch.equal(p.sorted_indices, orc.sort_plan(torch.tensor(g[f"len{n}"]))[1])


def test_tile_rows_heuristic():
    assert choose_tile_rows(1280, 148) == 128      # reference default batch 64: >= TC_MIN_SEQS -> 128-row tiles for the tensor-core GRU
    assert choose_tile_rows(20480, 148) == 128
    assert choose_tile_rows(320, 148) == 128       # the user->item side of batch 64
    assert choose_tile_rows(120, 148) == 32        # below two tiles: CUDA-core kernels


@pytest.mark.parametrize("sizes,ctas", [([160], 74), ([160, 160], 74), ([40, 160, 160], 74), ([5], 74), ([1], 1), ([300, 7], 3)])
def test_schedule_covers_every_tile_once_and_balances(sizes, ctas):
    from umpr_b200.plan import build_schedule
    rs = np.random.RandomState(sum(sizes))
    tile_lens = [np.sort(rs.randint(1, 21, size=n))[::-1] for n in sizes]
    sched, nq = build_schedule(tile_lens, ctas)
    T = sum(sizes)
    assert nq % 2 == 0 and nq // 2 == min(ctas, T) and sched.size == nq + 1 + T
    q_off, q_tile = sched[:nq + 1], sched[nq + 1:]
    assert q_off[0] == 0 and q_off[-1] == T and (np.diff(q_off) >= 0).all()
    assert sorted(q_tile.tolist()) == list(range(T))                    # every tile of every segment exactly once
    lens = np.concatenate(tile_lens)
    for q in range(nq):
        ql = lens[q_tile[q_off[q]:q_off[q + 1]]]
        assert (ql[:-1] >= ql[1:]).all()                                # longest first inside a queue
    per_q = np.array([lens[q_tile[q_off[q]:q_off[q + 1]]].sum() for q in range(nq)])
    if T >= nq:
        assert per_q.max() - per_q.min() <= 2 * lens.max()              # boustrophedon dealing keeps the slot queues level
    else:
        assert (per_q[0::2] > 0).all()                                  # few tiles: every CTA gets one before any gets two


def test_plan_prefetcher_prepares_the_same_plans_off_thread():
    """train.PlanPrefetcher (collate-worker pattern): plans built on the worker thread equal the ones built in place, ride on
    the lengths tensors of the unchanged 8-tuple, and every batch of the source comes out once, in order."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import PlanPrefetcher
    batches = [syn.make_batch("music_full", 9, vocab=500, seed=s) for s in range(4)]
    out = list(PlanPrefetcher(iter(batches), "cpu"))
    assert len(out) == 4
    for b, o in zip(batches, out):
        assert all(x is y for x, y in zip(b, o))
        for ids, lens in ((o[0], o[3]), (o[1], o[4]), (o[2], o[5])):
            ref = PackPlan(lens.reshape(-1), ids.shape[2], "cpu")
            got = lens._umpr_plan
            assert got.L == ids.shape[2] and got.N == lens.numel() and got.R == ref.R
            assert torch.equal(got.host, ref.host) and torch.equal(got.sorted_indices, ref.sorted_indices)
            assert got.ensure_uploaded().buf is not None


@pytest.mark.parametrize("L", [1, 20, 100, 128])
def test_snet_tile_table_covers_every_sentence_with_at_most_128_rows(L):
    rs = np.random.RandomState(L)
    lens = torch.from_numpy(rs.randint(1, L + 1, size=3000))
    p = PackPlan(lens, L, "cpu", tile_rows=128)
    tab, nt = p._snet_host()
    tso, cst = tab[:nt + 1], tab[nt + 1:]
    assert cst.size == p.N + 1 and tso[0] == 0 and tso[-1] == p.N and (np.diff(tso) >= 1).all()
    assert np.array_equal(np.diff(cst), p.row_lengths().numpy())            # per OUTPUT row, in output order
    rows = cst[tso[1:]] - cst[tso[:-1]]
    assert rows.max() <= 128 and int(rows.sum()) == p.tokens
    if L <= 20:
        assert rows[:-1].min() >= 129 - 2 * L                                # tiles are filled to within two sentences


@pytest.mark.parametrize("L", [1, 20, 100, 126])
def test_cnet_tile_table_has_guard_rows_and_at_most_128_rows(L):
    """The convolution's tile table: every sentence takes len + 2 tile rows (zero guard row before and after its valid rows)."""
    rs = np.random.RandomState(100 + L)
    lens = torch.from_numpy(rs.randint(1, L + 1, size=2500))
    p = PackPlan(lens, L, "cpu", tile_rows=128)
    tab, nt = p._cnet_host()
    tso, cst = tab[:nt + 1], tab[nt + 1:]
    assert cst.size == p.N + 1 and tso[0] == 0 and tso[-1] == p.N and (np.diff(tso) >= 1).all()
    assert np.array_equal(np.diff(cst), p.row_lengths().numpy() + 2)
    rows = cst[tso[1:]] - cst[tso[:-1]]
    assert rows.max() <= 128 and int(rows.sum()) == p.tokens + 2 * p.N
    assert np.diff(tso).max() <= 44                                          # CT_MAXS / CB_MAXS of the kernels: >= 3 rows per sentence


def test_max_valid_positions_per_sample():
    rs = np.random.RandomState(3)
    B, S, L = 50, 8, 30
    lens = torch.from_numpy(rs.randint(1, L + 1, size=B * S))
    p = PackPlan(lens, L, "cpu", tile_rows=128)
    eff = p.row_lengths().numpy().reshape(B, S)                              # lengths in OUTPUT-row order (the double un-sort)
    assert p.max_valid_per_sample(B) == int(eff.sum(1).max())


def test_prefetched_plan_uploads_plan_and_tables_in_one_buffer():
    """train.prepare_batch builds the pack plan and both valid-row tables off-thread; ``ensure_uploaded`` then ships them as ONE
    buffer and hands out the three views (plan, S-Net table, convolution table) the kernels read."""
    from umpr_b200 import synthetic as syn
    from umpr_b200.train import prepare_batch
    b = prepare_batch(syn.make_batch("music_full", 110, vocab=500, seed=1), "cpu")      # 2200 sentences per side: 128-row tiles
    lens = b[3]
    pl = lens._umpr_plan
    assert pl.R == 128 and pl.buf is None and pl._snet_np is not None and pl._cnet_np is not None
    pl.ensure_uploaded()
    assert torch.equal(pl.buf, pl.host)
    (st, s_nt), (ct, c_nt) = pl.snet_table(), pl.cnet_table()
    assert s_nt == pl._snet_np[1] and np.array_equal(st.numpy(), pl._snet_np[0])
    assert c_nt == pl._cnet_np[1] and np.array_equal(ct.numpy(), pl._cnet_np[0])
    assert st.is_contiguous() and ct.is_contiguous() and pl.buf.is_contiguous()
    # a side below the tensor-core threshold (2 * 5 = 10 user->item sentences): CUDA-core tiles, no tables are prepared for it
    tiny = prepare_batch(syn.make_batch("music_full", 2, vocab=500, seed=1), "cpu")[5]._umpr_plan
    assert tiny.R != 128 and getattr(tiny, "_snet_np", None) is None


@pytest.mark.parametrize("n,L", [(20480, 20), (5120, 20), (300, 33), (129, 128), (7, 5)])
def test_native_plan_builders_equal_the_numpy_specification(n, L):
    """csrc/plan_host.cu (umpr_plan_build / umpr_plan_table / umpr_plan_schedule) against the numpy forms in plan.py: identical pack
    plan, S-Net and convolution tile tables, and GRU tile schedules."""
    from umpr_b200 import plan as P
    rs = np.random.RandomState(n + L)
    lens = torch.from_numpy(rs.randint(1, L + 1, size=n).astype(np.int64))
    lens[torch.from_numpy(rs.rand(n) < 0.3)] = 1
    res = []
    for native in (True, False):
        P.NATIVE_PLAN = native
        try:
            p = P.PackPlan(lens, L, "cpu", tile_rows=128)
            res.append((p, p._snet_host(), p._cnet_host() if L <= 126 else None,
                        [P.build_schedule([p.tile_len, p.tile_len[::2]], c) for c in (1, 3, 74)]))
        finally:
            P.NATIVE_PLAN = True
    (a, ta, ca, sa), (b, tb, cb, sb) = res
    assert torch.equal(a.host, b.host) and (a.tokens, a.n_slabs, a.n_tiles) == (b.tokens, b.n_slabs, b.n_tiles)
    assert ta[1] == tb[1] and np.array_equal(ta[0], tb[0])
    if ca is not None:
        assert ca[1] == cb[1] and np.array_equal(ca[0], cb[0])
    for x, y in zip(sa, sb):
        assert x[1] == y[1] and np.array_equal(x[0], y[0])
