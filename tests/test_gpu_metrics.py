"""GPU parity tests of the calibration / uncertainty-error kernels, called through the Python drop-ins (ctypes ->
C-ABI -> CUDA).  Integer tables must be bit-exact against the oracle; float64 confidence sums may differ by
summation order only (rtol 1e-12, stated in SURVEY.md A4)."""
import numpy as np
import pytest
import torch

from rcu_b200 import evaluation as ev
from rcu_b200 import hooks, metrics, tables
from oracle import restate as R
from helpers import SWEEP, synth_metric_inputs, results_equal

pytestmark = pytest.mark.gpu
CONF_RTOL = 1e-12


def _check_calib(p, target, mask, n_subjects=1, **kw):
    cnt, pos, conf = metrics.calibration_tables(p, target, mask, n_subjects=n_subjects, **kw)
    vps = p.size // n_subjects
    for j in range(n_subjects):
        sl = slice(j * vps, (j + 1) * vps)
        c0, p0, s0 = R.calibration_tables(p[sl], target[sl], mask=None if mask is None else mask[sl],
                                          n_bins=kw.get('n_bins', 10), threshold_range=kw.get('threshold_range'))
        nb = kw.get('n_bins', 10)
        assert np.array_equal(cnt[j, :nb], c0[:nb]) and cnt[j, nb] == 0
        assert np.array_equal(pos[j, :nb], p0[:nb].astype(np.int64))
        assert np.allclose(conf[j, :nb], s0[:nb], rtol=CONF_RTOL, atol=0)


@pytest.mark.parametrize('n,s', [(1000003, 1), (155 * 240 * 240, 1), (4 * 65536, 4), (6 * 1001, 6), (37, 1), (4, 1), (1, 1)])
def test_calibration_tables_bit_exact(n, s):
    breaks, _, _ = tables.uncertainty_break_table()
    p, target, mask, _, _ = synth_metric_inputs(n, 20, with_break_neighbours=breaks)
    _check_calib(p, target, None, s)
    _check_calib(p, target, mask, s)


def test_calibration_other_bin_counts_and_range():
    p, target, mask, _, _ = synth_metric_inputs(300000, 21)
    for nb in (1, 7, 15, 32):
        _check_calib(p, target, mask, 1, n_bins=nb)
    _check_calib(p, target, None, 1, threshold_range=(0.2, 0.8))
    _check_calib(p, target, mask, 1, threshold_range=(0.0, 1.0))


def test_calibration_empty_and_all_masked():
    cnt, pos, conf = metrics.calibration_tables(np.zeros(0, np.float32), np.zeros(0, np.uint8))
    assert cnt.sum() == 0 and conf.sum() == 0
    p, target, _, _, _ = synth_metric_inputs(5000, 1)
    cnt, pos, conf = metrics.calibration_tables(p, target, np.zeros(5000, bool))
    assert cnt.sum() == 0 and pos.sum() == 0 and conf.sum() == 0


def test_calibration_out_of_range_policy():
    p = np.array([0.2, -0.1, 1.5, np.nan, 0.7, 1.0, 1.00000012], dtype=np.float32)
    cnt, _, _ = metrics.calibration_tables(p, np.zeros(7, np.uint8))
    assert cnt[0, 10] == 4 and cnt[0, :10].sum() == 3
    with pytest.raises(ValueError):
        ev.ece_binary(np.stack([1 - p, p], -1), np.zeros(7, np.uint8))


def test_conf_sum_is_deterministic_and_device_inputs_work():
    p, target, mask, _, _ = synth_metric_inputs(155 * 240 * 240, 2)
    a = metrics.calibration_tables(p, target, mask)
    pd, td, md = torch.from_numpy(p).cuda(), torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda()
    b = metrics.calibration_tables(pd, td, md)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def _check_ue(values, kind, pred, target, unc_ref, ths=SWEEP, mask=None, n_subjects=1):
    tab, inv, order = metrics.ue_tables(values, pred, target, ths, mask, kind=kind, n_subjects=n_subjects)
    assert inv.sum() == 0
    vps = pred.size // n_subjects
    sorted_ths = np.asarray(ths)[order]
    for j in range(n_subjects):
        sl = slice(j * vps, (j + 1) * vps)
        m = None if mask is None else mask[sl]
        assert tab[j].sum() == (vps if m is None else m.sum())
        for k, th in enumerate(sorted_ths):
            th = float(th)  # the reference compares against a Python float (weak scalar: float32 maps compare in float32)
            exp = R.uncertainty_counts(pred[sl].astype(bool), target[sl].astype(bool), unc_ref[sl] > th, m)
            assert tuple(int(v) for v in tables.counts_at_threshold(tab[j], k)) == tuple(int(v) for v in exp), (kind, j, th)


@pytest.mark.parametrize('n,s', [(1000003, 1), (155 * 240 * 240, 1), (4 * 65536, 4), (6 * 1001, 6), (37, 1)])
def test_ue_tables_bit_exact_all_kinds(n, s):
    breaks, _, _ = tables.uncertainty_break_table()
    p, target, mask, pred, _ = synth_metric_inputs(n, 20, with_break_neighbours=breaks)
    unc = R.normalized_entropy(R.add_background_probability(p))
    _check_ue(p, 'p', pred, target, unc, n_subjects=s)
    _check_ue(unc, 'u64', pred, target, unc, n_subjects=s)
    u32 = unc.astype(np.float32)
    _check_ue(u32, 'u32', pred, target, u32, n_subjects=s)
    _check_ue(p, 'p', pred, target, unc, mask=mask, n_subjects=s)


def test_ue_tables_unsorted_and_custom_thresholds():
    p, target, _, pred, _ = synth_metric_inputs(200000, 4)
    unc = R.normalized_entropy(R.add_background_probability(p))
    for ths in ((0.7, 0.1, 0.4), (0.5,), (-0.5, 1.5, 0.25)):
        _check_ue(p, 'p', pred, target, unc, ths)
        _check_ue(unc, 'u64', pred, target, unc, ths)


def test_ue_invalid_probabilities_are_counted():
    p = np.array([0.3, -0.2, 1.2, np.nan, 0.9], dtype=np.float32)
    z = np.zeros(5, np.uint8)
    _, inv, _ = metrics.ue_tables(p, z, z, kind='p')
    assert inv[0] == 3
    with pytest.raises(ValueError):
        ev.UncertaintySweepFromProbabilities()({'probabilities': p, 'prediction': z, 'target': z}, {})


@pytest.mark.parametrize('s', [1, 5])
def test_fused_equals_separate_and_confusion(s):
    n = s * 240 * 240 * 31
    p, target, mask, pred, _ = synth_metric_inputs(n, 6)
    c, po, cf, tab, inv, _ = metrics.eval_fused(p, pred, target, mask, n_subjects=s)
    c2, po2, cf2 = metrics.calibration_tables(p, target, mask, n_subjects=s)
    tab2, _, _ = metrics.ue_tables(p, pred, target, kind='p', n_subjects=s)
    assert np.array_equal(c, c2) and np.array_equal(po, po2) and np.array_equal(tab, tab2)
    # float64 confidence sums: the fused and the separate kernels split a subject into different block ranges, which
    # changes the summation order only
    assert np.allclose(cf, cf2, rtol=CONF_RTOL, atol=0)
    cm = metrics.confusion_counts(pred, target, n_subjects=s)
    assert np.array_equal(cm, tab.sum(axis=2))
    vps = n // s
    for j in range(s):
        assert tuple(cm[j]) == tuple(int(v) for v in R.confusion(pred[j * vps:(j + 1) * vps], target[j * vps:(j + 1) * vps])[:4])


def test_fifty_subjects_in_one_launch_match_per_subject_calls():
    s, vps = 50, 240 * 240 * 5
    p, target, mask, pred, _ = synth_metric_inputs(s * vps, 8)
    batched = metrics.eval_fused(p, pred, target, mask, n_subjects=s)
    for j in (0, 17, 49):
        sl = slice(j * vps, (j + 1) * vps)
        single = metrics.eval_fused(p[sl], pred[sl], target[sl], mask[sl])
        for i, (a, b) in enumerate(zip(batched[:5], single[:5])):
            if i == 2:  # float64 confidence sums: a different block partition changes the summation order only
                assert np.allclose(a[j], b[0], rtol=CONF_RTOL, atol=0)
            else:
                assert np.array_equal(a[j], b[0])


# ---------------------------------------------------------------------------------------------- drop-in strategies
def _golden_to_eval(g):
    return {'probabilities': g['prob2'], 'target': g['target'], 'prediction': g['prediction'], 'mask': g['mask'],
            'target_boarder': g['border'], 'uncertainty': g['uncertainty']}


@pytest.mark.parametrize('with_mask', [False, True])
@pytest.mark.parametrize('weighting', ['proportion', 'log_proportion', 'power_proportion', 'mean_proportion'])
def test_ece_strategy_matches_reference_golden(golden_metrics, with_mask, weighting):
    res = {}
    ev.EceBinaryNumpy(with_mask=with_mask, return_bins=True, bin_weighting=weighting)(_golden_to_eval(golden_metrics), res)
    pre = 'ece/mask%d/%s/' % (with_mask, weighting)
    assert sorted(res) == sorted(k.split('/')[-1] for k in golden_metrics.files if k.startswith(pre))
    for k, v in res.items():
        results_equal(v, golden_metrics[pre + k], k, rtol=0 if k in ('bins_count', 'bins_non_zero') else 1e-11)
    res = {}
    ev.EceBinaryNumpy(threshold_range=(0.2, 0.8))(_golden_to_eval(golden_metrics), res)
    results_equal(res['ece'], golden_metrics['ece/range/ece'], 'ece', rtol=1e-11)


def test_sweep_strategies_match_reference_golden(golden_metrics):
    to_eval = _golden_to_eval(golden_metrics)
    for th in SWEEP:
        res = {}
        ev.UncertaintyAndCorrectionEvalNumpy(th)(to_eval, res)
        keys = [k for k in golden_metrics.files if k.startswith('sweep/%s/' % th)]
        assert sorted(res) == sorted(k.split('/')[-1] for k in keys)
        for k in keys:
            results_equal(res[k.split('/')[-1]], golden_metrics[k], k)
        res = {}
        ev.UncertaintyErrorDiceNumpy(th, 'ue', with_mask=True)(to_eval, res)
        for k in ('precision', 'recall', 'dice'):
            results_equal(res['ue_' + k], golden_metrics['uedice_border/%s/ue_%s' % (th, k)], k)
        res = {}
        ev.UncertaintyErrorDiceNumpy(th)(to_eval, res)
        for k in ('precision', 'recall', 'dice'):
            results_equal(res[k], golden_metrics['uedice/%s/%s' % (th, k)], k)
    # the eleven sweep strategies shared one kernel pass
    assert sum(1 for k in to_eval['_rcu_b200'] if k[0] == 'u64') == 2   # one unmasked pass, one with the border mask
    res = {}
    ev.UncertaintySweepFromProbabilities()(to_eval, res)
    for th in SWEEP:
        for k in [k for k in golden_metrics.files if k.startswith('sweep/%s/' % th)]:
            results_equal(res['sweep'][th][k.split('/')[-1]], golden_metrics[k], k)
    res = {}
    ev.ComposeEvaluation([ev.DiceNumpy(), ev.ConfusionMatrix()])(to_eval, res)
    for k in ('dice', 'tp', 'tn', 'fp', 'fn', 'n'):
        results_equal(res[k], golden_metrics['dice_cm/' + k], k)


def test_degenerate_and_error_behaviour(golden_metrics):
    unc = golden_metrics['uncertainty']
    z = np.zeros(unc.shape, np.uint8)
    res = {}
    ev.UncertaintyAndCorrectionEvalNumpy(0.5)({'prediction': z, 'target': z, 'uncertainty': unc}, res)
    for k in [k for k in golden_metrics.files if k.startswith('degenerate_empty/0.5/')]:
        results_equal(res[k.split('/')[-1]], golden_metrics[k], k)
    with pytest.raises(ValueError, match='binary classification'):
        ev.ece_binary(np.zeros((5, 3), np.float32), np.zeros(5, np.uint8))
    with pytest.raises(ValueError, match="must be 'ndarray'"):
        ev.dice([1, 0], [1, 0])
    tp, tn, fp, fn, tpu, tnu, fpu, fnu = ev.uncertainty(np.array([1, 0, 1, 0]), np.array([1, 0, 0, 1]), np.array([True, False, True, False]))
    assert (tp, tn, fp, fn, tpu, tnu, fpu, fnu) == (1, 1, 1, 1, 1, 0, 1, 0)


def test_device_metrics_hook_rows():
    shape = (12, 64, 64)
    n = int(np.prod(shape))
    p, target, mask, pred, _ = synth_metric_inputs(n, 14)
    prob = np.stack([1 - p, p], -1).reshape(shape + (2,))
    pred = (prob[..., 1] > prob[..., 0]).astype(np.uint8)

    class Ctx:
        subject_index = 3
        subject_data = {'probabilities': prob, 'labels': target.reshape(shape), 'brain': mask.reshape(shape)}
    hook = hooks.DeviceMetricsHook(mask_entry='brain')
    hook.on_test_subject_end(Ctx, None, None)
    row = hook.rows[0]
    ece, bins = R.ece_binary(p.reshape(shape), target.reshape(shape), mask=mask.reshape(shape))
    assert np.isclose(row['ece'], ece, rtol=1e-11) and np.array_equal(row['bins_count'], bins['bins_count'])
    unc = R.normalized_entropy(R.add_background_probability(p.reshape(shape)))
    for th in SWEEP:
        exp = R.uncertainty_and_correction(pred, target.reshape(shape), unc, th)
        for k, v in exp.items():
            results_equal(row['sweep'][th][k], v, k)
        assert row['sweep'][th]['ue'] == R.ue_table_row(exp)['ue']
    assert row['dice'] == R.dice(pred, target.reshape(shape))


def test_both_fused_kernels_give_the_same_tables():
    """rcu_eval_fused has two kernels for aligned float32 input: the shared-atomic one (default; also with a 129-entry bucket table, which makes more buckets take the several-entries scan) and the
    bucket-table private-column one (RCU_HIST_ATOM=0).  The switches are read once per process, so the
    comparison runs in subprocesses: adversarial values at every bin edge and break point, NaN / negative /
    > 1 / inf, ragged sizes, several subjects, with and without mask — tables must be identical, the float64 confidence sums
    equal to 1e-12 (the overflow slot's sum is unspecified: the atomic kernel reports 0 there)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import rcu_b200
from rcu_b200 import metrics, tables
from helpers import synth_metric_inputs
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
p, target, mask, pred, _ = synth_metric_inputs(240007, 3, with_break_neighbours=bt[0])
p[100000:100012] = [np.nan, -0.5, 1.5, np.inf, -0.0, 1.0, 0.0, np.float32(1) - np.float32(2) ** -24, -1e-30, np.float32(1) + np.float32(2) ** -23, -np.inf, 1e-45]
out = []
for m in (mask, None):
    for n in (p.size, p.size - 3, 4 * 50000):
        r = metrics.eval_fused(p[:n], pred[:n], target[:n], None if m is None else m[:n], n_subjects=4 if n == 200000 else 1, break_table=bt)
        conf = np.array(r[2], dtype=np.float64)
        conf[:, -1] = 0.0
        out.append(np.concatenate([r[0].ravel(), r[1].ravel(), r[3].ravel(), r[4].ravel()]).astype(np.float64))
        out.append(np.nan_to_num(conf.ravel(), nan=-1.0))
np.save(sys.argv[1], np.concatenate(out))
''' % (root, os.path.join(root, 'tests'))
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for name, flags in (('atom', {}), ('lut', {'RCU_HIST_ATOM': '0'}), ('atom2', {'RCU_HIST_ATOM_BITS': '7'})):
            path = os.path.join(d, name + '.npy')
            env = dict(os.environ, **flags)
            r = subprocess.run([sys.executable, '-c', code, path], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            res[name] = np.load(path)
    for name in ('atom', 'atom2'):
        assert res['lut'].shape == res[name].shape
        assert np.allclose(res['lut'], res[name], rtol=1e-12, atol=0), name


def test_two_streams_of_one_device_do_not_share_a_workspace():
    """The kernels keep tickets and partial tables in a workspace; the wrapper holds one per (device, stream), so a background
    hook evaluating on a side stream next to the main loop gets the tables a lone call gets (several rounds in flight)."""
    n = 2 * 1000003
    p, target, mask, pred, _ = synth_metric_inputs(n, 31)
    p2, target2, mask2, pred2, _ = synth_metric_inputs(n, 32)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (p, pred, target, mask.view(np.uint8), p2, pred2, target2, mask2.view(np.uint8))]
    ref_a = [np.asarray(t) for t in metrics.eval_fused(dev[0], dev[1], dev[2], dev[3], 10, SWEEP)[:5]]
    ref_b = [np.asarray(t) for t in metrics.eval_fused(dev[4], dev[5], dev[6], dev[7], 10, SWEEP)[:5]]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    got_a, got_b = [], []
    for _ in range(8):
        got_a.append(metrics.eval_fused(dev[0], dev[1], dev[2], dev[3], 10, SWEEP, sync=False))
        with torch.cuda.stream(side):
            got_b.append(metrics.eval_fused(dev[4], dev[5], dev[6], dev[7], 10, SWEEP, sync=False))
    torch.cuda.synchronize()
    for res, ref in ((got_a, ref_a), (got_b, ref_b)):
        for r in res:
            for k in (0, 1, 3, 4):
                assert np.array_equal(np.asarray(r[k].cpu() if torch.is_tensor(r[k]) else r[k]), ref[k])
            conf = np.asarray(r[2].cpu() if torch.is_tensor(r[2]) else r[2])
            assert np.allclose(conf, ref[2], rtol=CONF_RTOL, atol=0)
