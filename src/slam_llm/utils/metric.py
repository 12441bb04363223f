"""Token accuracy (reference: src/slam_llm/utils/metric.py:3-20).  On the B200 step the same ratio is produced by
the fused cross-entropy kernel (n_correct / n_valid); this torch version serves recipes that call it directly."""
import torch


def compute_accuracy(pad_outputs, pad_targets, ignore_label):
    """pad_outputs [B, L] predicted ids, pad_targets [B, L] labels -> fraction correct over labels != ignore_label."""
    mask = pad_targets != ignore_label
    numerator = torch.sum(pad_outputs.masked_select(mask) == pad_targets.masked_select(mask))
    denominator = torch.sum(mask)
    return numerator.float() / denominator.float()
