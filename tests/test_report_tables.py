"""Report-side reductions: CSV rows of the eval hooks, data-set level ECE, best-threshold table.
tests/golden/tables_golden.npz holds what the UNMODIFIED reference wrote / computed for a small synthetic cohort
(tests/golden/make_golden_tables.py).  CPU tests pin the oracle and the host logic; the GPU test produces the same
CSV rows from the device tables."""
import numpy as np
import pytest

from rcu_b200 import tables
from oracle import restate as R
from helpers import SWEEP


def cohort(n_subjects=5, shape=(6, 30, 40), seed=4):   # mirrors tests/golden/make_golden_tables.py
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_subjects):
        p = rng.beta(0.3 + 0.1 * s, 0.3, size=shape).astype(np.float32)
        if s == 2:
            p = np.clip(p, 0.25, 1.0)
        target = (rng.random(shape) < p).astype(np.uint8)
        pred = (p > 0.5).astype(np.uint8)
        mask = rng.random(shape) < 0.5
        out.append((p, target, pred, mask))
    return out


def _parse(cell):
    if cell in ('True', 'False'):
        return cell == 'True'
    try:
        return int(cell)
    except ValueError:
        return float(cell)


def _rows_equal(header, rows, g_header, g_rows, rtol=0.0):
    assert list(header) == list(g_header)
    assert len(rows) == len(g_rows)
    for row, g_row in zip(rows, g_rows):
        for name, got, exp in zip(header, row, g_row):
            if name in ('test_id', 'subject_name'):
                assert got == exp
                continue
            exp = _parse(exp)
            if isinstance(exp, (bool, int)):
                assert type(exp)(got) == exp and not isinstance(got, float), (name, got, exp)
            elif np.isnan(exp):
                assert np.isnan(got), name
            else:
                assert np.isclose(float(got), exp, rtol=rtol, atol=0.0), (name, got, exp)


def _oracle_results():
    calib, sweeps = [], []
    for p, t, d, m in cohort():
        ece, bins = R.ece_from_tables(*R.calibration_tables(p, t, mask=m))
        res = dict(bins)            # return_bins entries first, then 'ece' (numpyfunctions.py:16-22)
        res['ece'] = ece
        calib.append(res)
        u = R.normalized_entropy(R.add_background_probability(p))
        sweeps.append({th: R.uncertainty_and_correction(d, t, u, th) for th in SWEEP})
    return calib, sweeps


def test_oracle_csv_rows_and_dataset_ece_match_reference(golden_tables):
    calib, sweeps = _oracle_results()
    rows = [R.bins_csv_row(r) for r in calib]
    header = ['test_id', 'subject_name'] + list(rows[0].keys())
    _rows_equal(header, [['baseline', 'subj%d' % i] + list(r.values()) for i, r in enumerate(rows)],
                golden_tables['calib/header'], golden_tables['calib/rows'])
    for th in SWEEP:
        rows_th = [s[th] for s in sweeps]
        _rows_equal(['test_id', 'subject_name'] + list(rows_th[0].keys()),
                    [['baseline', 'subj%d' % i] + list(r.values()) for i, r in enumerate(rows_th)],
                    golden_tables['ue/header'], golden_tables['ue/rows/%s' % th])
    ds = R.dataset_vs_mean_subject_ece(rows)
    # fp64 sums of ~10 terms in a different association than pandas / numpy.ma take them: 1e-12 relative
    assert np.isclose(ds['ece'], golden_tables['ds_ece/ece'], rtol=1e-12, atol=0)
    assert np.isclose(ds['ds_ece'], golden_tables['ds_ece/ds_ece'], rtol=1e-12, atol=0)
    assert not np.isclose(ds['ece'], ds['ds_ece'], rtol=0.05)      # the two notions really differ on this cohort


def test_host_reductions_follow_the_oracle(golden_tables):
    calib, sweeps = _oracle_results()
    header, rows = tables.csv_rows(calib, ['subj%d' % i for i in range(5)], 'baseline', bins=True)
    _rows_equal(header, rows, golden_tables['calib/header'], golden_tables['calib/rows'])
    assert len(calib[2]['bins_count']) < 10 and len(tables.expand_bins(calib[2])['bins_count']) == 10
    # data-set level ECE straight from the integer / fp64 tables
    tabs = [R.calibration_tables(p, t, mask=m) for p, t, _, m in cohort()]
    ds = tables.dataset_vs_mean_subject_ece(*(np.stack([tb[i] for tb in tabs]) for i in range(3)))
    assert np.isclose(ds['ece'], golden_tables['ds_ece/ece'], rtol=1e-12, atol=0)
    assert np.isclose(ds['ds_ece'], golden_tables['ds_ece/ds_ece'], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        tables.dataset_vs_mean_subject_ece(tabs[0][0], tabs[0][1], tabs[0][2])
    # best-threshold table against the restatement (itself pinned by test_best_threshold_reduction_matches_the_unmodified_reference_function)
    ece = [c['ece'] for c in calib]
    dice = [s[SWEEP[0]]['dice'] for s in sweeps]
    got = tables.best_threshold_summary(sweeps, ece, dice)
    exp = R.best_threshold_summary(sweeps, ece, dice)
    assert set(got) == set(exp)
    for k in exp:
        assert got[k] == exp[k], k
    assert got['benefit_threshold'] in SWEEP and got['error_threshold'] in SWEEP and 0 < got['error'] < 1
    # hand-made case: benefit only at one threshold, first maximum wins ties
    fake = [{0.1: dict(corrected_dice=0.5, dice=0.6, fn=1, fp=1, fnu=0, fpu=0, tnu=0, tpu=0),
             0.2: dict(corrected_dice=0.7, dice=0.6, fn=1, fp=1, fnu=1, fpu=1, tnu=0, tpu=0),
             0.3: dict(corrected_dice=0.7, dice=0.6, fn=1, fp=1, fnu=1, fpu=1, tnu=0, tpu=0)}]
    s = tables.best_threshold_summary(fake, [0.1], [0.6])
    assert s['benefit_threshold'] == 0.2 and s['benefit'] == 1.0 and s['error_threshold'] == 0.2 and s['error'] == 1.0


@pytest.mark.gpu
def test_device_hook_rows_become_the_reference_csv(golden_tables):
    import torch
    from rcu_b200 import hooks, metrics
    hook = hooks.DeviceMetricsHook()
    rows = []
    for i, (p, t, d, m) in enumerate(cohort()):
        rows.append(hook.evaluate('subj%d' % i, torch.from_numpy(p).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(t).cuda(),
                                  torch.from_numpy(m).cuda()))
    calib = [{k: r[k] for k in ('bins_count', 'bins_avg_confidence', 'bins_positive_fraction', 'bins_non_zero', 'ece')} for r in rows]
    header, csv = tables.csv_rows(calib, [r['subject'] for r in rows], 'baseline', bins=True)
    _rows_equal(header, csv, golden_tables['calib/header'], golden_tables['calib/rows'], rtol=1e-12)
    ue_entries = [e for e in golden_tables['ue/header'][2:]]
    for th in SWEEP:
        header, csv = tables.csv_rows([r['sweep'][th] for r in rows], [r['subject'] for r in rows], 'baseline', entries=ue_entries)
        _rows_equal(header, csv, golden_tables['ue/header'], golden_tables['ue/rows/%s' % th], rtol=1e-12)
    # batched tables of the whole cohort -> data-set level ECE
    stack = [np.concatenate([c[i].ravel() for c in cohort()]) for i in (0, 1, 3)]
    count, positives, conf = metrics.calibration_tables(stack[0], stack[1], stack[2], n_subjects=5)
    ds = tables.dataset_vs_mean_subject_ece(count[:, :10], positives[:, :10], conf[:, :10])
    assert np.isclose(ds['ece'], golden_tables['ds_ece/ece'], rtol=1e-12, atol=0)
    assert np.isclose(ds['ds_ece'], golden_tables['ds_ece/ds_ece'], rtol=1e-12, atol=0)
    summary = tables.best_threshold_summary([r['sweep'] for r in rows], [r['ece'] for r in rows], [r['dice'] for r in rows])
    _, sweeps = _oracle_results()
    exp = R.best_threshold_summary(sweeps, [r['ece'] for r in rows], [r['dice'] for r in rows])
    assert summary == exp


def test_best_threshold_reduction_matches_the_unmodified_reference_function():
    """tests/golden/best_golden.npz: `get_best_thresholds` (bin-analysis/table_ece_ue_bnf_dice.py:132-143), exec'd unmodified on
    a numeric frame, followed by the script's own merges and groupby('test_id').mean() — including a subject whose U-E
    Dice is 0 / 0 at two thresholds (pandas skips it; it must not disqualify those thresholds)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'best_golden.npz'))
    cols = [str(c) for c in g['columns']]
    ths = [float(t) for t in g['thresholds']]
    for ti in range(len(g['test_ids'])):
        d = g['data'][ti]                                   # [threshold][subject][column]
        sweeps = [{th: {c: (d[ki, s, ci] if c in ('corrected_dice', 'dice') else int(d[ki, s, ci])) for ci, c in enumerate(cols) if c != 'ece'}
                   for ki, th in enumerate(ths)} for s in range(d.shape[1])]
        ece, dice = d[0, :, cols.index('ece')], d[0, :, cols.index('dice')]
        exp = dict(zip([str(c) for c in g['result_columns']], g['result'][ti]))
        for impl in (tables.best_threshold_summary, R.best_threshold_summary):
            got = impl(sweeps, ece, dice, thresholds=ths)
            for k, v in exp.items():
                assert np.isclose(got[k], v, rtol=1e-12, atol=0), (impl.__module__, ti, k, got[k], v)
    assert np.isnan(g['data'][0, 1, 2, cols.index('fp')]) == False and 'groupby' in str(g['function_source'])
