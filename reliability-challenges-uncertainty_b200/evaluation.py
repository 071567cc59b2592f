"""GPU-backed drop-ins for the reference's metric layer.

Same class names, constructor arguments, `__call__(to_evaluate: dict, results: dict)` protocol, result keys and
error behaviour as common/evalutation/eval.py; same function names/signatures as
common/evalutation/numpyfunctions.py (the `np_fn` twins below).  Arrays under `to_evaluate` may be numpy arrays
(what bin-eval/eval_uncertainty.py builds) or CUDA tensors (the in-memory pipeline).

Differences worth knowing:
  * the eleven UncertaintyAndCorrectionEvalNumpy(th) objects of the CorrectionAction sweep share ONE kernel pass:
    the first call on a `to_evaluate` dict computes the joint table for all sweep thresholds and memoises it in
    the dict under '_rcu_b200';
  * a probability outside [0, 1+1e-8) (or NaN) raises ValueError — the reference either raises (negative values
    make np.bincount fail) or silently returns 11-long bin arrays (numpyfunctions.py:58-63).
"""
import abc

import numpy as np
import torch

from . import metrics
from . import tables

_CACHE_KEY = '_rcu_b200'


# --------------------------------------------------------------------------------------------------
# function-level twins of common/evalutation/numpyfunctions.py
# --------------------------------------------------------------------------------------------------
def _is_array(obj):
    return isinstance(obj, np.ndarray) or torch.is_tensor(obj)


def _check_ndarray(obj):
    if not _is_array(obj):
        raise ValueError("object of type '{}' must be '{}'".format(type(obj).__name__, np.ndarray.__name__))


def _foreground(probabilities, target):
    """binary_calibration's channel handling (numpyfunctions.py:27-33)."""
    if probabilities.ndim > target.ndim:
        if probabilities.shape[-1] > 2:
            raise ValueError('can only evaluate the calibration for binary classification')
        elif probabilities.shape[-1] == 2:
            probabilities = probabilities[..., 1]
        else:
            probabilities = probabilities.squeeze(-1) if torch.is_tensor(probabilities) else np.squeeze(probabilities, axis=-1)
    return probabilities


def _calibration_tables(probabilities, target, n_bins, threshold_range, mask):
    _check_ndarray(probabilities)
    _check_ndarray(target)
    p = _foreground(probabilities, target)
    count, positives, conf = metrics.calibration_tables(p, target, mask, n_bins, threshold_range)
    if count[0, n_bins] != 0:
        raise ValueError('{} probabilities are outside [0, 1+1e-8) or NaN'.format(int(count[0, n_bins])))
    return count[0, :n_bins], positives[0, :n_bins], conf[0, :n_bins]


def binary_calibration(probabilities, target, n_bins=10, threshold_range: tuple = None, mask=None):
    count, positives, conf = _calibration_tables(probabilities, target, n_bins, threshold_range, mask)
    nonzero = count != 0
    return positives[nonzero] / count[nonzero], conf[nonzero] / count[nonzero], count[nonzero], nonzero


def ece_binary(probabilities, target, n_bins=10, threshold_range: tuple = None, mask=None, out_bins: dict = None,
               bin_weighting='proportion'):
    count, positives, conf = _calibration_tables(probabilities, target, n_bins, threshold_range, mask)
    return tables.ece_from_tables(count, positives, conf, bin_weighting, target.ndim, out_bins)


def uncertainty(prediction, target, thresholded_uncertainty, mask=None):
    """tp, tn, fp, fn, tpu, tnu, fpu, fnu for a boolean thresholded-uncertainty map (numpyfunctions.py:86-107)."""
    u = thresholded_uncertainty
    u = u.to(torch.float32) if torch.is_tensor(u) else np.asarray(u, dtype=np.float32)
    table, _, _ = metrics.ue_tables(u, prediction, target, (0.5,), mask, kind='u32')
    return tables.counts_at_threshold(table[0], 0)


error_dice = tables.error_dice
error_recall = tables.error_recall
error_precision = tables.error_precision


def boarder_mask(binary_label_map, distance_in: int, distance_out: int):
    """common/utils/labelhelper.py:12-20 — the mask only (the summed distance map has no consumer on this path);
    numpy in -> numpy bool out, CUDA tensor in -> uint8 CUDA tensor out."""
    _check_ndarray(binary_label_map)
    mask = metrics.border_mask(binary_label_map, distance_in, distance_out)
    return mask if torch.is_tensor(binary_label_map) else mask.cpu().numpy().astype(bool)


def uncertainty_to_foreground_probabilities(uncertainty, prediction, rescale=None, epsilon=1e-5):
    """rechun/eval/helper.py:7-22 (optionally preceded by rescale_uncertainties, see metrics.confidence_to_foreground)."""
    _check_ndarray(uncertainty)
    _check_ndarray(prediction)
    out = metrics.confidence_to_foreground(uncertainty, prediction, rescale, epsilon)
    return out if torch.is_tensor(uncertainty) else out.cpu().numpy()


def confusion_matrx(prediction, target):
    _check_ndarray(prediction)
    _check_ndarray(target)
    tp, tn, fp, fn = metrics.confusion_counts(prediction, target)[0]
    return tp, tn, fp, fn, int(np.prod(prediction.shape))


def dice(prediction, target):
    tp, tn, fp, fn, _ = confusion_matrx(prediction, target)
    return tables.dice_from_counts(tp, fp, fn)


def accuracy(prediction, target):
    tp, tn, fp, fn, _ = confusion_matrx(prediction, target)
    return tables.accuracy_from_counts(tp, tn, fp, fn)


# --------------------------------------------------------------------------------------------------
# strategies (common/evalutation/eval.py)
# --------------------------------------------------------------------------------------------------
class EvaluationStrategy(metaclass=abc.ABCMeta):

    def __init__(self, result_entry=None) -> None:
        self.result_entry = result_entry

    @abc.abstractmethod
    def __call__(self, to_evaluate: dict, results: dict) -> None:
        pass


class EmptyEvaluation(EvaluationStrategy):
    def __call__(self, to_evaluate: dict, results: dict) -> None:
        pass


class ComposeEvaluation(EvaluationStrategy):

    def __init__(self, eval_strategies) -> None:
        super().__init__()
        self.eval_strategies = eval_strategies

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        for eval_ in self.eval_strategies:
            eval_(to_evaluate, results)


class LambdaEvaluation(EvaluationStrategy):

    def __init__(self, lambda_fn, entry_keys: tuple, result_entry) -> None:
        super().__init__(result_entry)
        self.lamda_fn = lambda_fn
        self.entry_keys = entry_keys

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        results[self.result_entry] = self.lamda_fn(*[to_evaluate[k] for k in self.entry_keys])


class NumpyEvaluationStrategy(EvaluationStrategy, metaclass=abc.ABCMeta):
    pass


class DiceNumpy(NumpyEvaluationStrategy):

    def __init__(self, result_entry='dice') -> None:
        super().__init__(result_entry)

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        results[self.result_entry] = dice(to_evaluate['prediction'], to_evaluate['target'])


class ConfusionMatrix(NumpyEvaluationStrategy):

    def __init__(self, result_entries=('tp', 'tn', 'fp', 'fn', 'n')) -> None:
        super().__init__(result_entries)

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        values = confusion_matrx(to_evaluate['prediction'], to_evaluate['target'])
        for key, value in zip(self.result_entry, values):
            results[key] = value


class EceBinaryNumpy(NumpyEvaluationStrategy):

    def __init__(self, n_bins=10, result_entry='ece', threshold_range: tuple = None, with_mask=False,
                 return_bins=False, bin_weighting='proportion') -> None:
        super().__init__(result_entry)
        self.n_bins = n_bins
        self.threshold_range = threshold_range
        self.with_mask = with_mask
        self.return_bins = return_bins
        self.bin_weighting = bin_weighting

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        mask = to_evaluate['mask'] if self.with_mask else None
        out_bins = results if self.return_bins else None
        results[self.result_entry] = ece_binary(to_evaluate['probabilities'], to_evaluate['target'], self.n_bins,
                                                self.threshold_range, mask, out_bins, self.bin_weighting)


def _as_bool_u8(x):
    # `.astype(np.bool)` of the reference (eval.py:159-160,184-185): any non-zero value is True
    if torch.is_tensor(x):
        return (x != 0).to(torch.uint8)
    return (np.asarray(x) != 0).view(np.uint8)


def _sweep_table(to_evaluate, threshold, mask, mask_src):
    """Joint table for `threshold`, shared between strategies evaluated on the same `to_evaluate` dict.

    The uncertainty entry decides the route: float64 numpy (what ToEntropy produces) -> exact float64 comparison on
    the device; float32 -> float32 break points.  Thresholds of the standard sweep are evaluated together and the
    table is memoised in the dict, so the eleven CorrectionAction strategies cost one kernel pass."""
    unc, prediction, target = to_evaluate['uncertainty'], to_evaluate['prediction'], to_evaluate['target']
    _check_ndarray(unc)
    _check_ndarray(prediction)
    _check_ndarray(target)
    is64 = (unc.dtype == np.float64) if isinstance(unc, np.ndarray) else (unc.dtype == torch.float64)
    kind = 'u64' if is64 else 'u32'
    group = tables.SWEEP_THRESHOLDS if threshold in tables.SWEEP_THRESHOLDS else (threshold,)
    cache = to_evaluate.setdefault(_CACHE_KEY, {})
    key = (kind, group, id(mask_src) if mask_src is not None else None, id(unc), id(prediction), id(target))
    if key not in cache:
        bkey = ('bool', id(prediction), id(target))
        if bkey not in cache:
            # the entries keep the keyed arrays alive: an id() can then never be recycled for another array of this dict
            cache[bkey] = (_as_bool_u8(prediction), _as_bool_u8(target), prediction, target)
        table, _, order = metrics.ue_tables(unc, cache[bkey][0], cache[bkey][1], group, mask, kind=kind)
        cache[key] = (table[0], list(np.asarray(group)[order]), unc, mask_src)
    table, sorted_ths = cache[key][:2]
    return table, sorted_ths.index(threshold)


class UncertaintyErrorDiceNumpy(NumpyEvaluationStrategy):

    def __init__(self, uncertainty_threshold, result_prefix: str = None, with_mask=False) -> None:
        super().__init__()
        self.uncertainty_threshold = uncertainty_threshold
        self.prefix = '' if result_prefix is None else result_prefix + '_'
        self.with_mask = with_mask

    def __call__(self, to_evaluate: dict, results: dict):
        mask = None
        if self.with_mask:
            boarder = to_evaluate['target_boarder']
            mask = (~boarder) if torch.is_tensor(boarder) else ~np.asarray(boarder, dtype=bool)
        table, k = _sweep_table(to_evaluate, self.uncertainty_threshold, mask, to_evaluate['target_boarder'] if self.with_mask else None)
        tp, tn, fp, fn, tpu, tnu, fpu, fnu = tables.counts_at_threshold(table, k)
        results['{}precision'.format(self.prefix)] = tables.error_precision(tpu, tnu, fpu, fnu)
        results['{}recall'.format(self.prefix)] = tables.error_recall(fp, fn, fpu, fnu)
        results['{}dice'.format(self.prefix)] = tables.error_dice(fp, fn, tpu, tnu, fpu, fnu)


class UncertaintyAndCorrectionEvalNumpy(NumpyEvaluationStrategy):

    def __init__(self, uncertainty_threshold) -> None:
        super().__init__()
        self.uncertainty_threshold = uncertainty_threshold

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        table, k = _sweep_table(to_evaluate, self.uncertainty_threshold, None, None)
        tables.correction_results(*tables.counts_at_threshold(table, k), results=results)


class UncertaintySweepFromProbabilities(NumpyEvaluationStrategy):
    """All sweep thresholds at once, straight from the saved foreground probability (no materialised uncertainty
    map): what CorrectionAction computes after AddBackgroundProbabilities + ToEntropy (bin-eval/eval_uncertainty.py:
    176-202, rechun/eval/analysis.py:255-258), 6 bytes per voxel instead of 11 numpy passes.

    results[result_entry] = {threshold: {...UncertaintyAndCorrectionEvalNumpy entries...}}"""

    def __init__(self, thresholds=tables.SWEEP_THRESHOLDS, result_entry='sweep') -> None:
        super().__init__(result_entry)
        self.thresholds = tuple(thresholds)
        self._table = tables.uncertainty_break_table(self.thresholds)

    def __call__(self, to_evaluate: dict, results: dict) -> None:
        probabilities = to_evaluate['probabilities']
        _check_ndarray(probabilities)
        p = _foreground(probabilities, to_evaluate['target'])
        table, invalid, order = metrics.ue_tables(p, _as_bool_u8(to_evaluate['prediction']), _as_bool_u8(to_evaluate['target']),
                                                  self.thresholds, None, kind='p', break_table=self._table)
        if invalid[0] != 0:
            # helper.add_background_probability -> check_min_max raises for p outside [0, 1] (rechun/eval/helper.py:25-47)
            raise ValueError('Found {} probabilities outside [0, 1]'.format(int(invalid[0])))
        out = {}
        for k_sorted, idx in enumerate(order):
            out[self.thresholds[idx]] = tables.correction_results(*tables.counts_at_threshold(table[0], k_sorted))
        results[self.result_entry] = out
