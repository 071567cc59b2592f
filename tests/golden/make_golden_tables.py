"""Generates tests/golden/tables_golden.npz: the report-side reductions of SURVEY.md §8f rank 3, produced by the
UNMODIFIED reference (through oracle/ref_shim.py): the CSV rows WriteCsvHook / WriteBinsCsvHook write for a small
synthetic cohort (rechun/eval/hook.py:27-93) and the data-set level vs mean-subject ECE the supplementary table
computes from the calibration CSV (bin-analysis/table_supplmat_ece_dataset_vs_meansubject.py:59-104).

    python tests/golden/make_golden_tables.py

`get_best_thresholds` (bin-analysis/table_ece_ue_bnf_dice.py:132-143) cannot be driven: under the installed pandas 3
its `groupby('run_id').mean()` over string columns raises TypeError (the reference pins pandas 0.25).  That reduction
is restated in oracle/restate.py and stays unpinned.
"""
import collections
import csv
import importlib.util
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, restate as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cohort(n_subjects=5, shape=(6, 30, 40), seed=4):
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_subjects):
        p = rng.beta(0.3 + 0.1 * s, 0.3, size=shape).astype(np.float32)
        if s == 2:
            p = np.clip(p, 0.25, 1.0)   # empty low bins: the compacted bins_* arrays are shorter than 10
        target = (rng.random(shape) < p).astype(np.uint8)
        pred = (p > 0.5).astype(np.uint8)
        mask = rng.random(shape) < 0.5
        out.append((p, target, pred, mask))
    return out


def read_csv(path):
    with open(path, newline='') as f:
        rows = list(csv.reader(f))
    return rows[0], rows[1:]


def main():
    ref_shim.load()
    warnings.simplefilter('ignore')
    import common.evalutation.eval as ev
    import rechun.eval.hook as hook
    import rechun.eval.analysis as analysis
    tmp = tempfile.mkdtemp()
    store = {}
    subjects = cohort()
    calib = hook.WriteBinsCsvHook(os.path.join(tmp, 'calib.csv'))
    ths = R.SWEEP_THRESHOLDS
    ue = {th: hook.WriteCsvHook(os.path.join(tmp, 'ue_%s.csv' % th)) for th in ths}
    for i, (p, t, d, m) in enumerate(subjects):
        te = {'probabilities': p.copy(), 'target': t, 'prediction': d, 'mask': m}
        te = analysis.ToEntropy()(analysis.AddBackgroundProbabilities()(te))
        res = {}
        ev.EceBinaryNumpy(with_mask=True, return_bins=True)(te, res)
        calib.on_subject(dict(res), 'subj%d' % i, 'baseline')
        for th in ths:
            r = {}
            ev.UncertaintyAndCorrectionEvalNumpy(th)(te, r)
            ue[th].on_subject(r, 'subj%d' % i, 'baseline')
    calib.on_run_end({}, 'baseline')
    header, rows = read_csv(os.path.join(tmp, 'calib.csv'))
    store['calib/header'] = np.array(header)
    store['calib/rows'] = np.array(rows)
    for th in ths:
        ue[th].on_run_end({}, 'baseline')
        header, rows = read_csv(os.path.join(tmp, 'ue_%s.csv' % th))
        store['ue/header'] = np.array(header)
        store['ue/rows/%s' % th] = np.array(rows)

    path = os.path.join(ref_shim.REFERENCE_ROOT, 'bin-analysis', 'table_supplmat_ece_dataset_vs_meansubject.py')
    spec = importlib.util.spec_from_file_location('refscript_ds_ece', path)
    table = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(table)
    table.get_brats_data = lambda: ([('baseline', os.path.join(tmp, 'calib.csv'))], collections.OrderedDict(baseline='baseline'))
    df = table.gather_information('brats')
    store['ds_ece/ece'] = np.float64(df.loc['baseline', 'ece'])
    store['ds_ece/ds_ece'] = np.float64(df.loc['baseline', 'ds_ece'])
    np.savez_compressed(os.path.join(OUT, 'tables_golden.npz'), **store)
    print('wrote tables_golden.npz', {k: (v.shape if hasattr(v, 'shape') else v) for k, v in store.items() if not k.startswith('ue/rows')})


if __name__ == '__main__':
    main()
