"""Fixture for the best-threshold reduction of bin-analysis/table_ece_ue_bnf_dice.py:30-73,132-143.

The script itself cannot be imported here (its module imports need the result folders), and `gather_information` reads
CSV files, so this generator
  * takes the source text of `get_best_thresholds` (lines 132-143) out of the reference file and exec()s it UNMODIFIED;
  * builds the frame the way `gather_information` does (:49-59: concat with run_id keys, `threshold`, `dice_diff`,
    `benefit`, `error`) from synthetic per-subject sweep results — with a NUMERIC `subject_name`, because under the
    installed pandas 3 `groupby.mean()` over a string column raises (the reference pins pandas 0.24, which dropped such
    nuisance columns silently); `test_id` is numeric for the same reason (0 = baseline_mc, 1 = ensemble);
  * applies the reference's own follow-up (:61-70: two merges, `groupby('test_id').mean()`).
One subject has no error voxel and no uncertain voxel at two thresholds (0/0 -> NaN `error`), which pins pandas' skipna
behaviour.  Run from the repo root in the authoring container:  python tests/golden/make_golden_best.py
"""
import ast
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get('RCU_REFERENCE', '/root/reference')
SRC = os.path.join(REF, 'bin-analysis', 'table_ece_ue_bnf_dice.py')


def reference_function(name):
    text = open(SRC).read()
    tree = ast.parse(text)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    src = '\n'.join(text.splitlines()[node.lineno - 1:node.end_lineno])
    ns = {'pd': pd}
    exec(compile(src, SRC, 'exec'), ns)
    return ns[name], src


def main():
    get_best_thresholds, src = reference_function('get_best_thresholds')
    rng = np.random.default_rng(7)
    thresholds = [0.05, 0.1, 0.2, 0.5, 0.9]
    test_ids = ['baseline_mc', 'ensemble']
    n_subjects = 6
    cols = ['corrected_dice', 'dice', 'fp', 'fn', 'fnu', 'fpu', 'tnu', 'tpu', 'ece']
    data = np.zeros((len(test_ids), len(thresholds), n_subjects, len(cols)))
    frames, run_ids = [], []
    for ti, test_id in enumerate(test_ids):
        dice = rng.random(n_subjects) * 0.5 + 0.4
        ece = rng.random(n_subjects) * 0.1
        for ki, th in enumerate(thresholds):
            fp, fn = rng.integers(0, 400, n_subjects), rng.integers(0, 300, n_subjects)
            fpu, fnu = (fp * rng.random(n_subjects) * (1 - th)).astype(np.int64), (fn * rng.random(n_subjects) * (1 - th)).astype(np.int64)
            tnu, tpu = rng.integers(0, 500, n_subjects), rng.integers(0, 200, n_subjects)
            corrected = dice + (rng.random(n_subjects) - 0.45) * 0.05
            if ti == 0 and ki in (1, 3):       # subject 2: nothing wrong and nothing uncertain -> error = 0 / 0
                for arr in (fp, fn, fpu, fnu, tnu, tpu):
                    arr[2] = 0
            data[ti, ki] = np.stack([corrected, dice, fp, fn, fnu, fpu, tnu, tpu, ece], axis=1)
            frame = pd.DataFrame({'test_id': ti, 'subject_name': np.arange(n_subjects), 'corrected_dice': corrected, 'fp': fp, 'fn': fn,
                                  'fnu': fnu, 'fpu': fpu, 'tnu': tnu, 'tpu': tpu, 'ece': ece, 'dice': dice})
            frames.append(frame)
            run_ids.append('{}_th{:03d}'.format(test_id, int(round(th * 100))))
    # gather_information, :49-70
    df = pd.concat(frames, keys=run_ids, names=['run_id'])
    ths = [float(s[-3:]) / 100 for s in list(df.index.get_level_values(0))]
    df['threshold'] = pd.Series(ths, index=df.index)
    df['dice_diff'] = df['corrected_dice'] - df['dice']
    df['benefit'] = df['dice_diff'] > 0
    df['error'] = (2 * (df['fnu'] + df['fpu'])) / (df['fn'] + df['fp'] + df['fnu'] + df['fpu'] + df['tnu'] + df['tpu'])
    assert df['error'].isna().sum() == 2
    best_benefit = get_best_thresholds(df[['test_id', 'subject_name', 'threshold', 'benefit']], 'benefit')
    best_benefit = best_benefit.rename(columns={'threshold': 'benefit_threshold'})
    best_error = get_best_thresholds(df[['test_id', 'subject_name', 'threshold', 'error']], 'error')
    best_error = best_error.rename(columns={'threshold': 'error_threshold'})
    out = df[['test_id', 'subject_name', 'ece', 'dice']]
    out = pd.merge(out, best_benefit, on=['test_id', 'subject_name'])
    out = pd.merge(out, best_error, on=['test_id', 'subject_name'])
    out = out.drop(columns='subject_name').groupby('test_id').mean()
    print(out)
    result = np.stack([out.loc[ti, ['ece', 'dice', 'benefit', 'benefit_threshold', 'error', 'error_threshold']].to_numpy(dtype=np.float64) for ti in range(len(test_ids))])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'best_golden.npz'), thresholds=np.array(thresholds), data=data,
                        columns=np.array(cols), test_ids=np.array(test_ids), result=result,
                        result_columns=np.array(['ece', 'dice', 'benefit', 'benefit_threshold', 'error', 'error_threshold']),
                        function_source=np.array(src))


if __name__ == '__main__':
    sys.exit(main())
