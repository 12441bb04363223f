"""Token accuracy (reference semantics: src/slam_llm/utils/metric.py:3-20).  On the B200 step the same ratio is produced by
the fused cross-entropy kernel (n_correct / n_valid); this torch version serves recipes that call it directly."""
import torch


def compute_accuracy(pad_outputs, pad_targets, ignore_label):
    """pad_outputs [B, L] predicted ids, pad_targets [B, L] labels -> fraction correct over labels != ignore_label (0-dim float tensor)."""
    counted = pad_targets.ne(ignore_label)
    hits = (pad_outputs.eq(pad_targets) & counted).sum()
    return hits.to(torch.float32) / counted.sum().to(torch.float32)
