"""Evaluation loop of the reference (src/evaluate.py:6-14): ``model.eval()``, ``no_grad`` forward over a loader of 8-tuples, sum of
squared errors divided by the sample count.  The forward is the same drop-in module on the CUDA kernels; the per-batch sum of
squares stays on the device and is read back once at the end (the reference calls ``.item()`` per batch, i.e. one sync per batch)."""
from __future__ import annotations

import torch


def evaluate_mse(model, dataloader) -> float:
    """evaluate.py:6-14.  ``dataloader`` yields the 8-tuples of dataset.py:173-182; returns mse over all samples."""
    was_training = model.training
    sq, count = None, 0
    with torch.no_grad():
        model.eval()
        for batch in dataloader:
            pred, _ = model(*batch)
            labels = batch[-1].to(pred.device, non_blocking=True).to(pred.dtype)
            e = ((pred - labels) ** 2).sum()                  # F.mse_loss(pred, labels, reduction='sum')
            sq = e if sq is None else sq + e
            count += pred.shape[0]
    if was_training:
        model.train()
    if count == 0:
        raise ZeroDivisionError("evaluate_mse: empty dataloader")      # the reference divides by sample_count == 0 here
    return float(sq) / count
