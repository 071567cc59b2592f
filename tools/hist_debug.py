"""The adversarial inputs of tests/test_gpu_metrics.py::test_all_three_fused_kernels_give_the_same_tables, run directly
(under compute-sanitizer when a kernel faults):  [compute-sanitizer] python tools/hist_debug.py"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import rcu_b200  # noqa
from rcu_b200 import metrics, tables
from helpers import synth_metric_inputs
import torch
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
p, target, mask, pred, _ = synth_metric_inputs(240007, 3, with_break_neighbours=bt[0])
p[100000:100012] = [np.nan, -0.5, 1.5, np.inf, -0.0, 1.0, 0.0, np.float32(1) - np.float32(2) ** -24, -1e-30, np.float32(1) + np.float32(2) ** -23, -np.inf, 1e-45]
for m in (mask, None):
    for n in (p.size, p.size - 3, 4 * 50000):
        r = metrics.eval_fused(p[:n], pred[:n], target[:n], None if m is None else m[:n], n_subjects=4 if n == 200000 else 1, break_table=bt)
        torch.cuda.synchronize()
        print(n, m is not None, r[0].sum(), r[3].sum(), r[4], flush=True)
