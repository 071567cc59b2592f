"""DeviceMetricsHook — evaluates each finished subject on the GPU inside the test loop, so the calibration and
uncertainty-error tables no longer need the NIfTI round trip through bin-eval/eval_uncertainty.py.

It follows the TestLoopHook callback protocol of common/trainloop/hooks.py:67-98 (only the callbacks it needs do
anything) and can be composed with the reference's hooks through ReducedComposeTestLoopHook (hooks.py:116-151).
Rows carry the entries the `ece_dice`, `calib` and `bnf_ue` actions write (bin-eval/eval_uncertainty.py:112-202).
"""
import numpy as np
import torch

from . import metrics
from . import tables


def subject_outputs(probabilities, prediction=None):
    """(foreground p float32, argmax prediction uint8) of an assembled subject — the two volumes WriteHook stores
    (bin-dl/brats_test_default.py:96-98; np.argmax ties -> class 0).  numpy in -> numpy out, tensor in -> tensor out."""
    # maps assembled from (N, 1, H, W) step outputs arrive channel-last with one channel
    if prediction is not None and prediction.ndim > 1 and prediction.shape[-1] == 1 and prediction.ndim == probabilities.ndim:
        prediction = prediction[..., 0]
    if probabilities.ndim > 1 and probabilities.shape[-1] == 1 and prediction is not None and probabilities.ndim == prediction.ndim + 1:
        probabilities = probabilities[..., 0]
    if torch.is_tensor(probabilities):
        if probabilities.shape[-1] == 2:
            return probabilities[..., 1].float().contiguous(), (probabilities[..., 1] > probabilities[..., 0]).to(torch.uint8)
        if prediction is None:
            raise ValueError('foreground probabilities need a "prediction" entry')
        return probabilities.float().contiguous(), (prediction != 0).to(torch.uint8)
    if probabilities.shape[-1] == 2:
        return (np.ascontiguousarray(probabilities[..., 1], dtype=np.float32),
                (probabilities[..., 1] > probabilities[..., 0]).astype(np.uint8))
    if prediction is None:
        raise ValueError('foreground probabilities need a "prediction" entry')
    return np.ascontiguousarray(probabilities, dtype=np.float32), np.asarray(prediction, dtype=np.uint8)


class DeviceMetricsHook:

    def __init__(self, thresholds=tables.SWEEP_THRESHOLDS, n_bins=10, mask_entry=None, label_entry='labels',
                 probability_entry='probabilities') -> None:
        self.thresholds = tuple(thresholds)
        self.n_bins = n_bins
        self.mask_entry = mask_entry
        self.label_entry = label_entry
        self.probability_entry = probability_entry
        self.break_table = tables.uncertainty_break_table(self.thresholds)
        self.rows = []

    # ---- callbacks that do nothing (protocol completeness) ----
    def on_startup(self): pass
    def end_startup(self, context): pass
    def on_termination(self, context): pass
    def on_test_start(self, task_context, context): pass
    def on_test_end(self, task_context, context): pass
    def on_test_batch_start(self, batch_context, task_context, context): pass
    def on_test_batch_end(self, batch_context, task_context, context): pass
    def on_test_subject_start(self, subject_context, task_context, context): pass

    def on_test_subject_end(self, subject_context, task_context, context):
        data = subject_context.subject_data
        prob = data[self.probability_entry]
        # what WriteHook saves and the eval script reloads (bin-dl/brats_test_default.py:96-98): argmax + foreground p.
        # Entries assembled by assembly.DeviceSubjectAssembler are CUDA tensors and never leave the device.
        p, prediction = subject_outputs(prob, data.get('prediction'))
        target = data[self.label_entry]
        target = (target != 0).to(torch.uint8) if torch.is_tensor(target) else (np.asarray(target) != 0).astype(np.uint8)
        mask = None
        if self.mask_entry is not None:
            mask = data[self.mask_entry]
            mask = (mask != 0) if torch.is_tensor(mask) else np.asarray(mask).astype(bool)
        self.rows.append(self.evaluate(subject_context.subject_index, p, prediction, target, mask))

    def evaluate(self, subject, p, prediction, target, mask=None):
        count, positives, conf, ue, invalid, order = metrics.eval_fused(p, prediction, target, mask, self.n_bins, self.thresholds,
                                                                        break_table=self.break_table)
        if invalid[0] or count[0, self.n_bins]:
            raise ValueError('subject {}: probabilities outside [0, 1]'.format(subject))
        row = {'subject': subject}
        bins = {}
        row['ece'] = tables.ece_from_tables(count[0, :self.n_bins], positives[0, :self.n_bins], conf[0, :self.n_bins],
                                            n_dim=target.ndim, out_bins=bins)
        row.update(bins)
        tp, tn, fp, fn = (ue[0, i].sum() for i in range(4))
        row.update(tp=tp, tn=tn, fp=fp, fn=fn, n=int(target.numel() if hasattr(target, 'numel') else np.size(target)), dice=tables.dice_from_counts(tp, fp, fn))
        row['sweep'] = {}
        for k_sorted, idx in enumerate(order):
            r = tables.correction_results(*tables.counts_at_threshold(ue[0], k_sorted))
            r.update(tables.ue_table_columns(r))
            row['sweep'][self.thresholds[idx]] = r
        return row
