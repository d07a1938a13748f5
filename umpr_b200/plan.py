"""Host-side length bookkeeping for ImprovedRnn (reference src/model.py:12-21).

The packing order is NOT re-implemented: it comes from the very call the reference makes
(``torch.sort(lengths.cpu(), descending=True)`` inside ``pack_padded_sequence(enforce_sorted=False)``,
model.py:18), because that sort is unstable and only the same call is bit-exact.
Everything derived from it is integer indexing.
"""
from __future__ import annotations

import torch

from . import _lib

TILE_ROWS = (32, 64, 128)


TC_MIN_SEQS = 2048     # from this many sequences on, a call runs on the fused tensor-core GRU kernel (tiles of 128)


def choose_tile_rows(n_seq: int, n_sm: int) -> int:
    """128-row tiles (the tcgen05 MMA's M) once there are enough sequences to feed the tensor-core kernel; below that the
    CUDA-core kernels with the largest tile whose (tile, direction) grid still covers every SM."""
    if n_seq >= TC_MIN_SEQS:
        return 128
    for r in (128, 64):
        if 2 * ((n_seq + r - 1) // r) >= n_sm:
            return r
    return 32


def build_schedule(tile_lens, n_ctas: int):
    """Tile queues of one fused GRU launch (umpr_gru_fwd_tc / umpr_gru_bwd_tc).

    ``tile_lens``: one int sequence per segment (steps of every 128-row tile, already descending inside a segment).
    Tiles of all segments get global ids (segment bases are cumulative tile counts) and are dealt longest-first, in
    boustrophedon order, to the 2*G slot queues of G = min(n_ctas, T) CTAs (queue order: slot 0 of every CTA, then slot 1 of
    every CTA, so few tiles spread over CTAs first).  A CTA's run time is its longer slot's step count, so it is the SLOT
    queues that are levelled; the two slots of a CTA ping-pong between tensor pipe and gate math.
    → (int32 tensor ``[q_off (2*G+1) | q_tile (T)]``, n_queues = 2*G); CTA c owns queues 2c and 2c+1.
    """
    import numpy as np
    lens = np.concatenate([np.asarray(t, dtype=np.int64).reshape(-1) for t in tile_lens])
    T = int(lens.size)
    if T == 0:
        raise RuntimeError("umpr_b200: empty GRU launch")
    G = max(1, min(int(n_ctas), T))
    Q = 2 * G
    order = np.argsort(-lens, kind="stable")
    i = np.arange(T)
    p, pos = i // Q, i % Q
    sq = np.where(p % 2 == 0, pos, Q - 1 - pos)             # slot queue in dealing order: [slot 0 of CTA 0..G-1 | slot 1 of CTA 0..G-1]
    queue = 2 * (sq % G) + sq // G                          # kernel order: CTA c reads queues 2c (slot 0) and 2c+1 (slot 1)
    by_q = np.argsort(queue, kind="stable")                 # stable: tiles stay longest-first inside a queue
    q_tile = order[by_q]
    counts = np.bincount(queue, minlength=Q)
    q_off = np.concatenate([[0], np.cumsum(counts)])
    return torch.from_numpy(np.concatenate([q_off, q_tile]).astype(np.int32)), 2 * G


class PackPlan:
    """Permutation, lengths and slab layout of one ImprovedRnn call.

    Attributes (host, int64): ``sorted_indices``, ``unsorted_indices`` — exactly those of the reference's
    PackedSequence; ``row_src[n] = unsorted_indices[n]`` is the sequence whose GRU output lands in result row ``n``
    (model.py:21 applies the un-sort permutation a second time).
    """

    def __init__(self, lengths: torch.Tensor, total_length: int, device, tile_rows: int | None = None):
        lens = lengths.detach().to("cpu", torch.int64).reshape(-1)          # model.py:18 lengths.cpu()
        n = lens.numel()
        if n == 0:
            raise RuntimeError("umpr_b200: ImprovedRnn needs at least one sequence")
        if int(lens.min()) < 1:
            # same failure as torch.nn.utils.rnn.pack_padded_sequence
            raise RuntimeError("Length of all samples has to be greater than 0, but found an element in 'lengths' that is <= 0")
        if int(lens.max()) > total_length:
            raise RuntimeError(f"umpr_b200: a sequence length ({int(lens.max())}) exceeds the padded length ({total_length})")
        sorted_len, sorted_idx = torch.sort(lens, descending=True)          # the reference's call, not a re-implementation
        unsorted = torch.empty_like(sorted_idx)
        unsorted[sorted_idx] = torch.arange(n, dtype=torch.int64)
        self.N, self.L = n, int(total_length)
        self.lengths = lens
        self.sorted_lengths, self.sorted_indices, self.unsorted_indices = sorted_len, sorted_idx, unsorted
        self.device = torch.device(device)
        self.R = tile_rows or choose_tile_rows(n, _lib.sm_count(self.device) if self.device.type == "cuda" else 148)
        assert self.R in TILE_ROWS
        R = self.R
        self.n_tiles = (n + R - 1) // R
        rp = self.n_tiles * R
        pad = rp - n
        z = torch.zeros(pad, dtype=torch.int64)
        seq_of = torch.cat([sorted_idx, z])
        row_of = torch.cat([sorted_idx[sorted_idx], z - 1])                  # output row fed by job k
        len_of = torch.cat([sorted_len, z])
        tile_len = len_of[::R]
        self.tile_len = tile_len
        tile_off = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(tile_len, 0)])
        self.n_slabs = int(tile_off[-1])
        slab_tile = torch.repeat_interleave(torch.arange(self.n_tiles, dtype=torch.int64), tile_len)
        host = torch.cat([seq_of, row_of, len_of, tile_off, slab_tile]).to(torch.int32)
        self.host = host
        self.tokens = int(lens.sum())                                        # T_v: valid tokens
        self.slots = self.n_slabs * R                                        # token slots actually computed
        if self.device.type == "cuda":
            self.buf = host.pin_memory().to(self.device, non_blocking=True)
        else:
            self.buf = host

    @property
    def row_src(self) -> torch.Tensor:
        return self.unsorted_indices

    def row_lengths(self) -> torch.Tensor:
        """Effective length of every OUTPUT row: len[unsorted_indices[n]] (the zero pattern of the result)."""
        return self.lengths[self.unsorted_indices]
