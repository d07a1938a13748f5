"""The data formats either side of the hot path.

``FeatureStore``  VGG16 feature cache: an on-disk table of the backbone's 1000-d output per photo (the reference runs VGG16 inside
                  ``VisualNet.forward`` on JPEGs loaded in ``batch_loader``, model.py:204-219, dataset.py:134-151; on this path the
                  backbone is upstream and ``photos`` are its features).  One file, memory-mapped; ``to_device`` makes it an
                  HBM-resident table (4 KB per photo: a million photos are 4 GB of the 180), and a batch then carries photo ROW
                  INDICES ``(B, V, Pc)`` that ``umpr_feature_gather`` turns into the ``(B, V, Pc, F)`` tensor the visual tail reads.
``collate``       the reference's ``batch_loader`` / ``pad_reviews`` (dataset.py:122-131,153-182) with the padding done ON THE DEVICE:
                  the host flattens the ragged sentence lists into one int32 token array + per-slot counts (numpy), ships those
                  (~2.4x fewer bytes than padded int64 tensors) and ``umpr_collate_ids`` expands them; lengths stay on the host as
                  the reference emits them (``max(1, len)``), because the packing order comes from the host ``torch.sort`` call.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

MAGIC = b"UMPRFEAT"
HEADER = 4096          # the feature matrix starts on a page boundary


class FeatureStore:
    """File layout: 4096-byte header ``[MAGIC | u32 version | u64 rows | u32 dim | u32 missing_row+1 | u32 json_bytes | json]`` where the
    JSON maps photo id -> row, then ``rows x dim`` little-endian fp32.  ``missing_row``: the row holding the backbone's features of
    the all-zero image the reference substitutes for an unreadable photo (dataset.py:147-148), or none (-> zeros)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(HEADER)
        if head[:8] != MAGIC:
            raise RuntimeError(f"{path}: not a umpr_b200 feature store")
        version, rows, dim, miss, jlen = struct.unpack_from("<IQIII", head, 8)
        if version != 1:
            raise RuntimeError(f"{path}: feature store version {version} not supported")
        self.rows, self.dim, self.missing_row = int(rows), int(dim), int(miss) - 1
        off = 8 + struct.calcsize("<IQIII")
        if jlen and off + jlen <= HEADER:
            self.index = json.loads(head[off:off + jlen].decode())
        else:                                           # large key tables live in a side file
            with open(path + ".keys.json") as f:
                self.index = json.load(f)
        self.features = np.memmap(path, dtype="<f4", mode="r", offset=HEADER, shape=(self.rows, self.dim))
        self.table = None                               # device copy (to_device)

    @staticmethod
    def build(path, photo_ids, features, missing_features=None):
        """Write a store: ``features[i]`` (fp32, dim) belongs to ``photo_ids[i]``; ``missing_features``: the backbone's output for the
        zero image (appended as the last row)."""
        feats = np.ascontiguousarray(np.asarray(features, dtype="<f4"))
        if feats.ndim != 2 or feats.shape[0] != len(photo_ids):
            raise ValueError("features must be (len(photo_ids), dim)")
        miss = 0
        if missing_features is not None:
            feats = np.concatenate([feats, np.asarray(missing_features, dtype="<f4").reshape(1, -1)], 0)
            miss = feats.shape[0]                        # stored as row + 1
        index = {str(k): i for i, k in enumerate(photo_ids)}
        if len(index) != len(photo_ids):
            raise ValueError("duplicate photo ids")
        js = json.dumps(index).encode()
        fixed = 8 + struct.calcsize("<IQIII")
        inline = fixed + len(js) <= HEADER
        head = bytearray(HEADER)
        head[:8] = MAGIC
        struct.pack_into("<IQIII", head, 8, 1, feats.shape[0], feats.shape[1], miss, len(js) if inline else 0)
        if inline:
            head[fixed:fixed + len(js)] = js
        with open(path, "wb") as f:
            f.write(head)
            f.write(feats.tobytes())
        if not inline:
            with open(path + ".keys.json", "w") as f:
                f.write(js.decode())
        return FeatureStore(path)

    def rows_of(self, photo_ids):
        """Photo ids (any nesting, e.g. ``(B, V, Pc)`` lists; the reference pads short views with 'unknown', dataset.py:113-115) ->
        int32 row indices, -1 for ids the store does not hold."""
        a = np.asarray(photo_ids, dtype=object)
        flat = np.fromiter((self.index.get(str(k), -1) for k in a.reshape(-1)), dtype=np.int32, count=a.size)
        return flat.reshape(a.shape)

    def to_device(self, device):
        self.table = torch.from_numpy(np.array(self.features, dtype=np.float32, copy=True)).to(device)
        return self

    def gather(self, rows, device=None):
        """``rows`` (B, V, Pc) int -> (B, V, Pc, dim) fp32 features on the device (``umpr_feature_gather``)."""
        if self.table is None:
            self.to_device(device if device is not None else "cuda")
        dev = self.table.device
        idx = torch.as_tensor(np.asarray(rows), dtype=torch.int32)
        idx = (idx.pin_memory() if idx.device.type == "cpu" and dev.type == "cuda" else idx).to(dev, non_blocking=True).contiguous()
        out = torch.empty(*idx.shape, self.dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("umpr_feature_gather", ptr(self.table), ptr(idx), idx.numel(), self.rows, self.dim, self.missing_row, ptr(out))
        return out


def _flatten(reviews, max_count):
    """Ragged ``[sample][sentence][token]`` lists -> (flat int32 tokens, int32 counts (B, max_count)); slots beyond a sample's
    sentences are empty (dataset.py:125)."""
    B = len(reviews)
    counts = np.zeros((B, max_count), dtype=np.int32)
    for b, sents in enumerate(reviews):
        counts[b, :len(sents)] = [len(s) for s in sents]
    flat = np.fromiter((t for sents in reviews for s in sents for t in s), dtype=np.int32, count=int(counts.sum()))
    return flat, counts


def collate(batch_list, device, *, feature_store: FeatureStore = None, ignore_photos: bool = False, pad: int = 0):
    """The reference's ``batch_loader`` (dataset.py:153-182) for samples ``(user_sents, item_sents, ui_sents, photo_ids, rating)``:
    → the 8-tuple ``UMPR.forward`` takes, with the three id tensors ALREADY padded on ``device`` (int64, as the reference emits),
    the lengths on the host (int64, ``max(1, len)``), and ``photos`` = features gathered from ``feature_store`` by photo id
    (``Tensor([])`` when ``ignore_photos``, dataset.py:158,180)."""
    dev = torch.device(device)
    data = [[s[i] for s in batch_list] for i in range(3)]
    # user and item share (max_count, max_len) (dataset.py:163-170); ui is padded to its own maxima (:171)
    mc = max(max(len(ru), len(ri)) for ru, ri in zip(data[0], data[1]))
    ml = max(max(max(len(s) for s in ru), max(len(s) for s in ri)) for ru, ri in zip(data[0], data[1]))
    out_ids, out_len = [], []
    for k in range(3):
        count = mc if k < 2 else max(len(s) for s in data[2])
        flat, counts = _flatten(data[k], count)
        lengths = np.maximum(counts, 1).astype(np.int64)                     # dataset.py:127
        L = ml if k < 2 else int(lengths.max())
        off = np.zeros(counts.size + 1, dtype=np.int32)
        np.cumsum(counts.reshape(-1), out=off[1:])
        host = torch.from_numpy(np.concatenate([off, flat]))
        host = host.pin_memory() if dev.type == "cuda" else host
        buf = host.to(dev, non_blocking=True)
        ids = torch.empty(len(batch_list), count, L, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            call("umpr_collate_ids", ptr(buf[off.size:]) if flat.size else None, ptr(buf[:off.size]), counts.size, L, pad, ptr(ids))
        out_ids.append(ids)
        out_len.append(torch.from_numpy(lengths))
    if ignore_photos:
        photos = torch.zeros(0)
    else:
        if feature_store is None:
            raise RuntimeError("umpr_b200: collate needs a FeatureStore for the photos (or ignore_photos=True)")
        photos = feature_store.gather(feature_store.rows_of([s[3] for s in batch_list]), dev)
    labels = torch.tensor([s[4] for s in batch_list], dtype=torch.float32)
    return (*out_ids, *out_len, photos, labels)
