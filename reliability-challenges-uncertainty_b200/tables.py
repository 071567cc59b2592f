"""Host-side logic of the calibration / uncertainty-error metrics: the small parameter tables the CUDA kernels
consume, and the reductions that turn the kernels' integer tables into the reference's result entries.

Nothing here touches the GPU; everything is covered by the CPU test-suite.

Reference arithmetic these tables encode (path:line relative to the reference root):
  np.linspace(0., 1. + 1e-8, n_bins + 1) + np.digitize      common/evalutation/numpyfunctions.py:53-54
  u(p) = -sum(where(q > 0, q*log(q), [0.0])) / log(2)       common/evalutation/numpyfunctions.py:166-168,
         over q in (fl32(1-p), p)                           rechun/eval/helper.py:25-28, rechun/eval/analysis.py:201
  u > threshold                                             common/evalutation/eval.py:167,188
"""
import math

import warnings

import numpy as np

SWEEP_THRESHOLDS = (0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95)  # bin-eval/eval_uncertainty.py:239
MAX_BREAKS = 96
_ONE_BITS = 0x3F800000
_HALF_BITS = 0x3F000000


# --------------------------------------------------------------------------------------------------
# ECE bin edges
# --------------------------------------------------------------------------------------------------
def calibration_edges_f32(n_bins=10):
    """Smallest float32 >= each float64 edge of np.linspace(0, 1+1e-8, n_bins+1).

    For a float32 p and a float64 edge e:  p >= e  <=>  p >= ceil32(e), so np.digitize against the float64
    edges (numpyfunctions.py:53-54) is reproduced exactly by float32 comparisons on the device.
    """
    edges64 = np.linspace(0., 1. + 1e-8, n_bins + 1)
    e32 = edges64.astype(np.float32)
    low = e32.astype(np.float64) < edges64
    e32[low] = np.nextafter(e32[low], np.float32(np.inf))
    return e32


# --------------------------------------------------------------------------------------------------
# p -> "number of sweep thresholds exceeded" break table
# --------------------------------------------------------------------------------------------------
def reference_uncertainty_of_p(p32):
    """The eval-side uncertainty of a saved float32 foreground probability, in the reference's arithmetic
    (float32 products, float64 sum and division) — helper.add_background_probability + np_fn.entropy / ln 2."""
    p32 = np.asarray(p32, dtype=np.float32)
    prob = np.stack([1 - p32, p32], axis=-1)
    with np.errstate(divide='ignore', invalid='ignore'):
        return -np.where(prob > 0, prob * np.log(prob), [0.0]).sum(axis=-1) / np.log(2)


def _accurate_uncertainty_of_bits(bits):
    """u for the float32 with bit pattern `bits`, evaluated in float64 from the same fl32(1-p), p pair."""
    p32 = np.asarray(bits, dtype=np.uint32).view(np.float32)
    q = (np.float32(1) - p32).astype(np.float64)
    p = p32.astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        h = -(np.where(q > 0, q * np.log(q), 0.0) + np.where(p > 0, p * np.log(p), 0.0))
    return h / math.log(2)


def _bisect_bits(lo, hi, pred_true_at_hi):
    """First bit pattern in (lo, hi] where the monotone predicate becomes true (true at hi, false at lo)."""
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if pred_true_at_hi(mid):
            hi = mid
        else:
            lo = mid
    return hi


def _zone(th, rel_eps, rising):
    """Bit range [a, b] on one side of 0.5 where |u_accurate - th| <= eps: the only floats whose float32-product
    uncertainty can land on the other side of `th` than the accurate value does."""
    eps = rel_eps * max(abs(th), 1e-30)
    if rising:   # p in [0, 0.5]: u increases with p
        lo, hi = 0, _HALF_BITS
        a = _bisect_bits(lo - 1, hi + 1, lambda b: b > hi or _accurate_uncertainty_of_bits(b) >= th - eps)
        b = _bisect_bits(lo - 1, hi + 1, lambda b: b > hi or _accurate_uncertainty_of_bits(b) > th + eps) - 1
    else:        # p in [0.5, 1]: u decreases with p
        lo, hi = _HALF_BITS, _ONE_BITS
        a = _bisect_bits(lo - 1, hi + 1, lambda b: b > hi or _accurate_uncertainty_of_bits(b) <= th + eps)
        b = _bisect_bits(lo - 1, hi + 1, lambda b: b > hi or _accurate_uncertainty_of_bits(b) < th - eps) - 1
    return max(a, lo), min(b, hi)


def uncertainty_break_table(thresholds=SWEEP_THRESHOLDS, rel_eps=4e-6, max_zone=1 << 25):
    """Break table for classifying a float32 foreground probability by how many thresholds its reference-arithmetic
    uncertainty exceeds — exactly, for every float32 in [0, 1].

    Returns (breaks float32[B] ascending, seg_class uint8[B+1], order) where a probability p with
    i = #{b : breaks[b] <= p} exceeds exactly seg_class[i] of the ascending-sorted thresholds, and `order` is the
    argsort that maps sorted positions back to the caller's threshold order.

    u(p) in float32-product arithmetic is not monotone near a threshold crossing (rounding noise of a few 1e-8
    against a slope that vanishes towards p = 0.5), so each crossing is resolved by brute force: all floats whose
    accurately evaluated uncertainty lies within a relative `rel_eps` of the threshold (the rounding noise is below
    1e-6 relative) are evaluated with the reference expression and every flip becomes a break point.  Outside those
    zones the reference value and the accurate value are on the same side of the threshold.
    (tools/ue_interval_exhaustive.py checks the same thing over all 2^30 floats for the default sweep.)
    """
    ths = np.asarray(thresholds, dtype=np.float64)
    if ths.ndim != 1 or ths.size == 0:
        raise ValueError('thresholds must be a non-empty 1-d sequence')
    if np.isnan(ths).any():
        raise ValueError('thresholds must not be NaN')
    order = np.argsort(ths, kind='stable')
    sorted_ths = ths[order]
    flips = set()
    for th in sorted_ths:
        if th < 0:
            continue  # u >= 0 > th for every p in [0, 1]: exceeded everywhere, no break
        for rising in (True, False):
            a, b = _zone(float(th), rel_eps, rising)
            if b < a:
                # no float is near the threshold on this side; the crossing (if any) is a clean step
                lo, hi = (0, _HALF_BITS) if rising else (_HALF_BITS, _ONE_BITS)
                acc = _accurate_uncertainty_of_bits(np.array([lo, hi], dtype=np.uint32)) > th
                if acc[0] != acc[1]:
                    if rising:
                        flips.add(_bisect_bits(lo, hi, lambda x: _accurate_uncertainty_of_bits(x) > th))
                    else:
                        flips.add(_bisect_bits(lo, hi, lambda x: not (_accurate_uncertainty_of_bits(x) > th)))
                continue
            if b - a + 1 > max_zone:
                raise ValueError('threshold {} is too close to a flat part of the entropy curve to tabulate'.format(th))
            lo_edge = 0 if rising else _HALF_BITS
            hi_edge = _HALF_BITS if rising else _ONE_BITS
            a2, b2 = max(a - 1, lo_edge), min(b + 1, hi_edge)
            bits = np.arange(a2, b2 + 1, dtype=np.int64).astype(np.uint32)
            on = reference_uncertainty_of_p(bits.view(np.float32)) > th
            # the floats just outside the zone are "certain": force them to the accurate value so that a zone
            # clipped at 0, 0.5 or 1 still yields the right boundary flips
            change = np.flatnonzero(on[1:] != on[:-1]) + 1
            for i in change:
                flips.add(int(bits[i]))
    flips = sorted(flips)
    if len(flips) > MAX_BREAKS:
        raise ValueError('{} break points exceed the kernel table size {}'.format(len(flips), MAX_BREAKS))
    breaks = np.asarray(flips, dtype=np.uint32).view(np.float32) if flips else np.zeros(0, dtype=np.float32)
    # class of each segment = thresholds exceeded by its first float (segments are constant by construction)
    starts = np.asarray([0] + flips, dtype=np.uint32).view(np.float32)
    u = reference_uncertainty_of_p(starts)
    seg_class = (u[:, None] > sorted_ths[None, :]).sum(axis=1).astype(np.uint8)
    return breaks, seg_class, order


def threshold_breaks_f32(thresholds):
    """For a float32 uncertainty map numpy evaluates `u > th` in float32 (the Python-float threshold is cast to the
    array's dtype, under value-based casting as well as NEP 50):  u > fl32(th)  <=>  u >= nextafter(fl32(th), +inf).
    Returns (breaks, identity seg_class, order)."""
    ths = np.asarray(thresholds, dtype=np.float64)
    order = np.argsort(ths, kind='stable')
    b = np.nextafter(ths[order].astype(np.float32), np.float32(np.inf))
    return b, np.arange(len(ths) + 1, dtype=np.uint8), order


# --------------------------------------------------------------------------------------------------
# Reductions of the kernels' tables into the reference's result entries
# --------------------------------------------------------------------------------------------------
def ece_from_tables(count, positives, conf_sum, bin_weighting='proportion', n_dim=3, out_bins=None):
    """ece_binary's tail (numpyfunctions.py:6-23) + _binary_calibration's compaction (:65-69) + _get_proportion (:72-83).

    count/positives/conf_sum are the n_bins-long tables (without the overflow slot)."""
    bin_total = np.asarray(count, dtype=np.int64)
    bin_true = np.asarray(positives).astype(np.float64)
    bin_sums = np.asarray(conf_sum, dtype=np.float64)
    nonzero = bin_total != 0
    pos_frac = bin_true[nonzero] / bin_total[nonzero]
    mean_confidence = bin_sums[nonzero] / bin_total[nonzero]
    bin_count = bin_total[nonzero]
    if bin_weighting == 'proportion':
        bin_proportions = bin_count / bin_count.sum()
    elif bin_weighting == 'log_proportion':
        bin_proportions = np.log(bin_count) / np.log(bin_count).sum()
    elif bin_weighting == 'power_proportion':
        bin_proportions = bin_count ** (1 / n_dim) / (bin_count ** (1 / n_dim)).sum()
    elif bin_weighting == 'mean_proportion':
        bin_proportions = 1 / nonzero.sum()
    else:
        raise ValueError('unknown bin weighting "{}"'.format(bin_weighting))
    if out_bins is not None:
        out_bins['bins_count'] = bin_count
        out_bins['bins_avg_confidence'] = mean_confidence
        out_bins['bins_positive_fraction'] = pos_frac
        out_bins['bins_non_zero'] = nonzero
    return (np.abs(mean_confidence - pos_frac) * bin_proportions).sum()


def counts_at_threshold(ue_counts, k_sorted):
    """tp, tn, fp, fn, tpu, tnu, fpu, fnu (np.int64) for the k-th smallest threshold from a [4][K+1] joint table."""
    t = np.asarray(ue_counts, dtype=np.int64)
    plain = t.sum(axis=1)
    with_u = t[:, k_sorted + 1:].sum(axis=1)
    return tuple(plain) + tuple(with_u)


def dice_from_counts(tp, fp, fn):
    # pymia 0.2.1 DiceCoefficient.calculate
    if tp == 0 and (tp + fp + fn) == 0:
        return 1.
    return 2 * tp / (2 * tp + fp + fn)


def accuracy_from_counts(tp, tn, fp, fn):
    # pymia 0.2.1 Accuracy.calculate
    s = tp + tn + fp + fn
    return (tp + tn) / s if s != 0 else 0


def error_dice(fp, fn, tpu, tnu, fpu, fnu):
    if (fnu + fpu) == 0 and (fn + fp + fnu + fpu + tnu + tpu) == 0:
        return 1.
    return (2 * (fnu + fpu)) / (fn + fp + fnu + fpu + tnu + tpu)


def error_recall(fp, fn, fpu, fnu):
    if (fnu + fpu) == 0 and (fn + fp) == 0:
        return 1.
    return (fnu + fpu) / (fn + fp)


def error_precision(tpu, tnu, fpu, fnu):
    if (fnu + fpu) == 0 and (fnu + fpu + tpu + tnu) == 0:
        return 1.
    return (fnu + fpu) / (fnu + fpu + tpu + tnu)


def correction_results(tp, tn, fp, fn, tpu, tnu, fpu, fnu, results=None):
    """Every entry UncertaintyAndCorrectionEvalNumpy writes (eval.py:192-226), from the eight counts alone.

    Correcting the thresholded-uncertain voxels to background moves tpu from tp to fn and fpu from fp to tn;
    correcting them to foreground moves fnu from fn to tp and tnu from tn to fp."""
    r = {} if results is None else results
    # numpy integers like the reference's np.sum results: 0 / 0 and x / 0 give nan / inf (compared below), not an exception
    tp, tn, fp, fn, tpu, tnu, fpu, fnu = (np.int64(v) for v in (tp, tn, fp, fn, tpu, tnu, fpu, fnu))
    r['tpu'], r['tnu'], r['fpu'], r['fnu'] = tpu, tnu, fpu, fnu
    r['tp'], r['tn'], r['fp'], r['fn'] = tp, tn, fp, fn
    with np.errstate(divide='ignore', invalid='ignore'):
        tpu_fpu_ratio = r['tpu'] / r['fpu']
        jaccard_index = r['tp'] / (r['tp'] + r['fp'] + r['fn'])
    r['dice_benefit'] = tpu_fpu_ratio < jaccard_index
    r['accuracy_benefit'] = tpu_fpu_ratio < 1
    r['dice'] = dice_from_counts(tp, fp, fn)
    r['accuracy'] = accuracy_from_counts(tp, tn, fp, fn)
    r['corrected_dice'] = dice_from_counts(tp - tpu, fp - fpu, fn + tpu)
    r['corrected_accuracy'] = accuracy_from_counts(tp - tpu, tn + fpu, fp - fpu, fn + tpu)
    r['dice_benefit_correct'] = (r['corrected_dice'] > r['dice']) == r['dice_benefit']
    r['accuracy_benefit_correct'] = (r['corrected_accuracy'] > r['accuracy']) == r['accuracy_benefit']
    r['corrected_add_dice'] = dice_from_counts(tp + fnu, fp + tnu, fn - fnu)
    r['corrected_add_accuracy'] = accuracy_from_counts(tp + fnu, tn - tnu, fp + tnu, fn - fnu)
    return r


def ue_table_columns(r):
    """Per-subject derived columns of bin-analysis/table_ece_ue_bnf_dice.py:56-59."""
    denom = r['fn'] + r['fp'] + r['fnu'] + r['fpu'] + r['tnu'] + r['tpu']
    with np.errstate(divide='ignore', invalid='ignore'):
        ue = (2 * (r['fnu'] + r['fpu'])) / np.float64(denom)
    return {'benefit': r['corrected_dice'] > r['dice'], 'ue': ue}


# --------------------------------------------------------------------------------------------------
# Report-side reductions: what the eval hooks write and what the analysis tables compute from it
# --------------------------------------------------------------------------------------------------
def expand_bins(results):
    """WriteBinsCsvHook.on_subject (rechun/eval/hook.py:75-93): the compacted `bins_*` arrays of `return_bins=True`
    re-expanded to full length with zeros in the empty bins (`bins_non_zero` says which)."""
    out = dict(results)
    non_zero = np.asarray(results['bins_non_zero'], dtype=bool)
    for key in ('bins_count', 'bins_avg_confidence', 'bins_positive_fraction'):
        value = np.asarray(results[key])
        full = np.zeros(non_zero.shape, dtype=value.dtype)
        full[non_zero] = value
        out[key] = full
    return out


def unfold_results(results):
    """WriteCsvHook._unfold_results (rechun/eval/hook.py:49-62): sequences become `key_00 ... key_NN` columns."""
    unfolded = {}
    for key, value in results.items():
        if isinstance(value, np.ndarray):
            value = value.tolist()
        if isinstance(value, (list, tuple)):
            nb_digits = len(str(len(value)))
            for i in range(len(value)):
                unfolded['{}_{:0{}d}'.format(key, i, nb_digits)] = value[i]
        else:
            unfolded[key] = value
    return unfolded


def csv_rows(results_per_subject, subject_names, run_id, entries=None, bins=False):
    """(header, rows) exactly as WriteCsvHook / WriteBinsCsvHook (hook.py:27-47,75-93) would write them."""
    header, rows = None, []
    for results, name in zip(results_per_subject, subject_names):
        unfolded = unfold_results(expand_bins(results) if bins else results)
        if entries is None:
            entries = list(unfolded.keys())
        if header is None:
            header = ['test_id', 'subject_name'] + list(entries)
        rows.append([run_id, name] + [unfolded[e] for e in entries])
    return header, rows


def dataset_vs_mean_subject_ece(count, positives, conf_sum):
    """bin-analysis/table_supplmat_ece_dataset_vs_meansubject.py:59-86 from the per-subject tables themselves
    ((S, n_bins) arrays, e.g. the batched output of rcu_calib_hist / rcu_eval_fused): the mean of the per-subject
    ECEs and the ECE of the pooled data set (bins summed over subjects before the ratio)."""
    count = np.asarray(count, dtype=np.int64)
    positives = np.asarray(positives, dtype=np.int64)
    conf_sum = np.asarray(conf_sum, dtype=np.float64)
    if count.ndim != 2 or count.shape != positives.shape or count.shape != conf_sum.shape:
        raise ValueError('expected three (subjects, n_bins) tables')
    ece = np.array([ece_from_tables(c, p, s) for c, p, s in zip(count, positives, conf_sum)])
    return {'ece': ece.mean(), 'ds_ece': ece_from_tables(count.sum(0), positives.sum(0), conf_sum.sum(0))}


def best_threshold_summary(sweeps, ece, dice, thresholds=None):
    """The row bin-analysis/table_ece_ue_bnf_dice.py:30-73 prints for one test id.

    sweeps: per subject, {threshold: UncertaintyAndCorrectionEvalNumpy results} (what DeviceMetricsHook rows hold under
    'sweep'); ece / dice: per-subject values.  `benefit` = corrected_dice > dice, `error` = the U-E Dice (:56-59); for
    each of the two the threshold with the best subject-mean is chosen (`get_best_thresholds`, :132-143; first maximum
    in threshold order) and the subject-mean at that threshold reported.  Both means are pandas means: a subject whose U-E
    Dice is 0 / 0 at a threshold (nothing wrong, nothing uncertain) is left out of that threshold's mean (skipna), it does
    not make the threshold ineligible.  Pinned by tests/golden/best_golden.npz (the unmodified `get_best_thresholds`)."""
    if thresholds is None:
        thresholds = list(sweeps[0].keys())
    benefit = np.array([[float(s[th]['corrected_dice'] - s[th]['dice'] > 0) for s in sweeps] for th in thresholds])
    error = np.array([[ue_table_columns(s[th])['ue'] for s in sweeps] for th in thresholds], dtype=np.float64)

    def best(table):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)      # all-NaN threshold -> NaN, like DataFrame.mean
            means = np.nanmean(table, axis=1)                    # per threshold over subjects, NaN subjects skipped
        if np.all(np.isnan(means)):
            return float('nan'), float('nan')
        k = int(np.nanargmax(means))                             # Series.idxmax: first maximum, NaN skipped
        return float(means[k]), float(thresholds[k])
    b_mean, b_th = best(benefit)
    e_mean, e_th = best(error)
    return {'ece': float(np.mean(ece)), 'dice': float(np.mean(dice)), 'benefit': b_mean, 'benefit_threshold': b_th,
            'error': e_mean, 'error_threshold': e_th}
