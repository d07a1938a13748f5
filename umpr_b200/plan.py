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


TC_MIN_SEQS = 128      # from this many sequences on (one full tile), a call runs on the fused tensor-core GRU kernel (tiles of 128):
                       # the reference default batch_size=64 (1280 sentences per side, 320 user->item sentences) is on tcgen05


def choose_tile_rows(n_seq: int, n_sm: int) -> int:
    """128-row tiles (the tcgen05 MMA's M) once there are enough sequences to feed the tensor-core kernel; below that the
    CUDA-core kernels with the largest tile whose (tile, direction) grid still covers every SM."""
    if n_seq >= TC_MIN_SEQS:
        return 128
    for r in (128, 64):
        if 2 * ((n_seq + r - 1) // r) >= n_sm:
            return r
    return 32


SMS_RESERVED_FOR_COMM = 0   # SMs the R-Net GRU launches leave free (train.FlatTrainer sets it when the gradient all-reduce overlaps the
                            # backward): the fused GRU kernels are persistent with one CTA per SM and nearly all of its shared memory, so
                            # an NCCL kernel that starts beside them would otherwise take SMs away from their grid for as long as it waits
                            # for the slowest rank
NATIVE_PLAN = True     # integer bookkeeping in C (csrc/plan_host.cu) when the library is there; False = the numpy forms below (the specification)


def _native_lib():
    if not NATIVE_PLAN:
        return None
    try:
        return _lib.load()
    except RuntimeError:
        return None


def build_schedule(tile_lens, n_ctas: int):
    """Tile queues of one fused GRU launch - ``umpr_plan_schedule`` (C) or ``build_schedule_np``; see the latter for the layout."""
    lib = _native_lib()
    if lib is None:
        return build_schedule_np(tile_lens, n_ctas)
    import ctypes as C
    import numpy as np
    arrs = [np.ascontiguousarray(t, dtype=np.int64).reshape(-1) for t in tile_lens]
    T = sum(a.size for a in arrs)
    if T == 0:
        raise RuntimeError("umpr_b200: empty GRU launch")
    G = max(1, min(int(n_ctas), T))
    out = np.empty(2 * G + 1 + T, dtype=np.int32)
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    nts = (C.c_int32 * len(arrs))(*[a.size for a in arrs])
    nq = C.c_int32(0)
    if lib.umpr_plan_schedule(ptrs, nts, len(arrs), int(n_ctas), out.ctypes.data, out.size, C.byref(nq)) != 0:
        raise RuntimeError(f"umpr_plan_schedule: {_lib.last_error()}")
    return out, nq.value


def build_schedule_np(tile_lens, n_ctas: int):
    """Tile queues of one fused GRU launch (umpr_gru_fwd_tc / umpr_gru_bwd_tc).

    ``tile_lens``: one int sequence per segment (steps of every 128-row tile, already descending inside a segment).
    Tiles of all segments get global ids (segment bases are cumulative tile counts) and are dealt longest-first, in
    boustrophedon order, to the 2*G slot queues of G = min(n_ctas, T) CTAs (queue order: slot 0 of every CTA, then slot 1 of
    every CTA, so few tiles spread over CTAs first).  A CTA's run time is its longer slot's step count, so it is the SLOT
    queues that are levelled; the two slots of a CTA ping-pong between tensor pipe and gate math.
    → (int32 numpy array ``[q_off (2*G+1) | q_tile (T)]``, n_queues = 2*G); CTA c owns queues 2c and 2c+1.
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
    return np.concatenate([q_off, q_tile]).astype(np.int32), 2 * G


class _PinnedRing:
    """Small ring of pinned host buffers for the per-step int32 plan / schedule uploads (cudaHostAlloc per step is far too
    slow, and a pageable source can stall the host behind queued GPU work).  A slot is reused only after the copy that last
    read it has completed (event)."""

    def __init__(self, slots=32, slot_ints=1 << 17):
        # one pinned allocation up front: cudaHostAlloc inside the training loop costs milliseconds and synchronises
        self.slots = slots
        self._alloc(slot_ints)

    def _alloc(self, slot_ints):
        self.slot_ints = slot_ints
        self.pool = torch.empty(self.slots * slot_ints, dtype=torch.int32).pin_memory()
        self.bufs = [self.pool[i * slot_ints:(i + 1) * slot_ints] for i in range(self.slots)]
        self.events, self.i = [None] * self.slots, 0

    def upload(self, arr, device):
        import numpy as np
        n = int(arr.size)
        if n > self.slot_ints:
            # a plan larger than a slot (batches of several thousand samples): ONE re-allocation of the whole ring with slots of
            # twice the size needed - not one cudaHostAlloc per slot as the ring rotates.  In-flight copies finish first.
            for e in self.events:
                if e is not None:
                    e.synchronize()
            self._alloc(1 << (2 * n - 1).bit_length())
        i = self.i
        self.i = (i + 1) % self.slots
        if self.events[i] is not None:
            self.events[i].synchronize()
        host = self.bufs[i][:n]
        host.numpy()[:] = np.asarray(arr, dtype=np.int32).reshape(-1)
        dev = host.to(device, non_blocking=True)
        if self.events[i] is None:
            self.events[i] = torch.cuda.Event()
        self.events[i].record()
        return dev


_RINGS = {}


def upload_int32(arr, device):
    """int32 numpy array → device tensor through the pinned ring of that device (plain copy for CPU 'devices')."""
    device = torch.device(device)
    if device.type != "cuda":
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32))
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _RINGS:
        _RINGS[key] = _PinnedRing()
    return _RINGS[key].upload(arr, device)


def _tile_starts(window, n):
    """First sentence of every tile (+ n at the end) for a non-decreasing window index per sentence: a tile starts wherever the
    index changes (what ``searchsorted`` over ``np.unique(...)`` ids gives, without the sort)."""
    import numpy as np
    change = np.empty(n, dtype=bool)
    change[0] = True
    np.not_equal(window[1:], window[:-1], out=change[1:])
    return np.append(np.flatnonzero(change), n)


class PackPlan:
    """Permutation, lengths and slab layout of one ImprovedRnn call.

    Attributes (host, int64): ``sorted_indices``, ``unsorted_indices`` — exactly those of the reference's
    PackedSequence; ``row_src[n] = unsorted_indices[n]`` is the sequence whose GRU output lands in result row ``n``
    (model.py:21 applies the un-sort permutation a second time).
    """

    def __init__(self, lengths: torch.Tensor, total_length: int, device, tile_rows: int | None = None, upload: bool = True):
        import numpy as np
        lens = lengths.detach().to("cpu", torch.int64).reshape(-1)          # model.py:18 lengths.cpu()
        n = lens.numel()
        if n == 0:
            raise RuntimeError("umpr_b200: ImprovedRnn needs at least one sequence")
        sorted_len, sorted_idx = torch.sort(lens, descending=True)          # the reference's call, not a re-implementation
        if int(sorted_len[-1]) < 1:
            # same failure as torch.nn.utils.rnn.pack_padded_sequence
            raise RuntimeError("Length of all samples has to be greater than 0, but found an element in 'lengths' that is <= 0")
        if int(sorted_len[0]) > total_length:
            raise RuntimeError(f"umpr_b200: a sequence length ({int(sorted_len[0])}) exceeds the padded length ({total_length})")
        self.N, self.L = n, int(total_length)
        self.lengths = lens
        self.sorted_lengths, self.sorted_indices = sorted_len, sorted_idx
        self._unsorted = None
        self.device = torch.device(device)
        self.R = tile_rows or choose_tile_rows(n, _lib.sm_count(self.device) if self.device.type == "cuda" else 148)
        assert self.R in TILE_ROWS
        R = self.R
        self.n_tiles = (n + R - 1) // R
        rp = self.n_tiles * R
        # everything below is integer indexing of the reference's permutation, done in numpy (this runs every step)
        si, sl = sorted_idx.numpy(), sorted_len.numpy()
        tile_len = np.zeros(self.n_tiles, dtype=np.int64)
        tile_len[:] = sl[::R]
        self.tile_len = tile_len
        self.n_slabs = int(tile_len.sum())
        host = np.empty(3 * rp + self.n_tiles + 1 + self.n_slabs, dtype=np.int32)
        lib = _native_lib()
        if lib is not None:
            import ctypes as C
            tokens = C.c_int64(0)
            if lib.umpr_plan_build(si.ctypes.data, sl.ctypes.data, n, R, host.ctypes.data, host.size, C.byref(tokens)) != 0:
                raise RuntimeError(f"umpr_plan_build: {_lib.last_error()}")
            self._finish(host, int(tokens.value), upload)
            return
        host[:n] = si                                                        # seq_of
        host[n:rp] = 0
        host[rp:rp + n] = si[si]                                             # row_of: output row fed by job k (model.py:21)
        host[rp + n:2 * rp] = -1
        host[2 * rp:2 * rp + n] = sl                                         # len_of
        host[2 * rp + n:3 * rp] = 0
        host[3 * rp] = 0
        np.cumsum(tile_len, out=host[3 * rp + 1:3 * rp + 1 + self.n_tiles])  # tile_off
        host[3 * rp + 1 + self.n_tiles:] = np.repeat(np.arange(self.n_tiles, dtype=np.int32), tile_len)   # slab_tile
        self._finish(host, int(sl.sum()), upload)

    def _finish(self, host, tokens, upload):
        self.host = torch.from_numpy(host)
        self.tokens = tokens                                                 # T_v: valid tokens
        self.slots = self.n_slabs * self.R                                   # token slots actually computed
        self._host_np = host
        self.buf = upload_int32(host, self.device) if upload else None      # upload=False: built off-thread, see ensure_uploaded()

    def ensure_uploaded(self):
        """Plans prepared by a data-pipeline thread (train.PlanPrefetcher) are uploaded here, on the consumer's thread/stream."""
        if self.buf is None:
            # one H2D copy for the plan and whichever valid-row tables the worker has already built
            import numpy as np
            parts, names = [self._host_np], []
            for name in ("_snet", "_cnet"):
                t = getattr(self, name + "_np", None)
                if t is not None and getattr(self, name, None) is None:
                    parts.append(t[0])
                    names.append((name, t[1]))
            dev = upload_int32(np.concatenate(parts) if len(parts) > 1 else parts[0], self.device)
            o = self._host_np.size
            self.buf = dev[:o]
            for (name, nt), part in zip(names, parts[1:]):
                setattr(self, name, (dev[o:o + part.size], nt))
                o += part.size
        return self

    @property
    def unsorted_indices(self) -> torch.Tensor:
        if self._unsorted is None:
            u = torch.empty_like(self.sorted_indices)
            u[self.sorted_indices] = torch.arange(self.N, dtype=torch.int64)
            self._unsorted = u
        return self._unsorted

    @property
    def row_src(self) -> torch.Tensor:
        return self.unsorted_indices

    def snet_table(self):
        """Tile table of the tensor-core S-Net (csrc/snet_tc.cu): consecutive OUTPUT rows (sentences, in the order ImprovedRnn
        returns them) are grouped so that a tile holds at most 128 valid positions.  Sentence n goes to tile
        ``cstart[n] // (129 - L)``: starts inside a window of 129-L rows plus one sentence of at most L rows never exceed 128.
        → (device int32 ``[tile_sent_off (n_tiles+1) | cstart (N+1)]``, n_tiles); cached."""
        if getattr(self, "_snet", None) is None:
            self._snet = (upload_int32(self._snet_host()[0], self.device), self._snet_host()[1])
        return self._snet

    def max_valid_per_sample(self, B: int) -> int:
        """Largest number of valid positions any of the ``B`` samples has (its sentences are consecutive output rows)."""
        key = "_pvmax%d" % B
        if getattr(self, key, None) is None:
            cs = self._snet_host()[0][self._snet_host()[1] + 1:]
            S = self.N // B
            setattr(self, key, int((cs[S::S] - cs[:-1:S]).max()))
        return getattr(self, key)

    def _table_native(self, extra):
        """[tile_sent_off | cstart] by ``umpr_plan_table`` → (int32 array, n_tiles), or None without the library."""
        lib = _native_lib()
        if lib is None:
            return None
        import ctypes as C
        import numpy as np
        buf = np.empty(2 * (self.N + 1), dtype=np.int32)
        nt = C.c_int32(0)
        si, ln = self.sorted_indices.numpy(), self.lengths.numpy()
        if lib.umpr_plan_table(si.ctypes.data, ln.ctypes.data, self.N, self.L, extra, buf.ctypes.data, C.byref(nt)) != 0:
            raise RuntimeError(f"umpr_plan_table: {_lib.last_error()}")
        return buf[:nt.value + 1 + self.N + 1], nt.value

    def _snet_host(self):
        if getattr(self, "_snet_np", None) is None:
            import numpy as np
            if self.L > 128:
                raise RuntimeError(f"umpr_b200: S-Net sentence length {self.L} exceeds 128")
            nat = self._table_native(0)
            if nat is not None:
                self._snet_np = nat
                return self._snet_np
            n = self.N
            row_len = np.empty(n, dtype=np.int64)
            # output row row_of[k] = si[si[k]] holds sequence si[k] (model.py:21), i.e. row si[j] holds sequence j
            row_len[self.sorted_indices.numpy()] = self.lengths.numpy()
            cstart = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(row_len, out=cstart[1:])
            tso = _tile_starts(cstart[:-1] // (129 - self.L), n)       # a new tile wherever the window index changes: no empty tiles
            nt = tso.size - 1
            self._snet_np = (np.concatenate([tso, cstart]).astype(np.int32), nt)
        return self._snet_np

    def cnet_table(self):
        """Tile table of the tensor-core C-Net convolution (csrc/cnet_tc.cu): like ``snet_table`` with every sentence taking
        ``len + 2`` tile rows (a zero guard row before and after its valid rows).  → (device int32 ``[tile_sent_off | cstart]``,
        n_tiles); cached."""
        if getattr(self, "_cnet", None) is None:
            host, nt = self._cnet_host()
            self._cnet = (upload_int32(host, self.device), nt)
        return self._cnet

    def _cnet_host(self):
        if getattr(self, "_cnet_np", None) is None:
            import numpy as np
            if self.L + 2 > 128:
                raise RuntimeError(f"umpr_b200: C-Net sentence length {self.L} exceeds 126")
            nat = self._table_native(2)
            if nat is not None:
                self._cnet_np = nat
                return self._cnet_np
            n = self.N
            stab, s_nt = self._snet_host()                                   # prefix sums of the lengths are already there
            cstart = stab[s_nt + 1:].astype(np.int64) + 2 * np.arange(n + 1, dtype=np.int64)
            tso = _tile_starts(cstart[:-1] // (129 - (self.L + 2)), n)
            nt = tso.size - 1
            self._cnet_np = (np.concatenate([tso, cstart]).astype(np.int32), nt)
        return self._cnet_np

    def row_lengths(self) -> torch.Tensor:
        """Effective length of every OUTPUT row: len[unsorted_indices[n]] (the zero pattern of the result)."""
        return self.lengths[self.unsorted_indices]
