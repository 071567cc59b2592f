"""Data dependence of the fused histogram kernels: device time for one 8.9 M-voxel subject on iid maps, on constant maps
(every voxel in one cell) and on smooth maps (runs of equal cells), per kernel (RCU_HIST_ATOM=0/1 in the environment)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa
from rcu_b200 import metrics, tables
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
n = 155 * 240 * 240
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(3)


def timeit(fn, reps=12):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


cases = {}
p = torch.rand(n, device=dev, generator=g)
cases['iid uniform p, Bernoulli target, 25% mask'] = (p, (p > 0.5).to(torch.uint8), (torch.rand(n, device=dev, generator=g) < p).to(torch.uint8),
                                                      (torch.rand(n, device=dev, generator=g) < 0.25).to(torch.uint8))
z = torch.zeros(n, dtype=torch.uint8, device=dev)
cases['constant p = 0.01, all labels 0, mask 1'] = (torch.full((n,), 0.01, device=dev), z, z, torch.ones(n, dtype=torch.uint8, device=dev))
cases['constant p = 0.01, all labels 0, mask 0'] = (torch.full((n,), 0.01, device=dev), z, z, z)
x = torch.linspace(0, 40 * np.pi, n, device=dev)
ps = torch.sigmoid(3 * torch.sin(x) + 0.3 * torch.sin(17.3 * x))
cases['smooth p (long runs of equal cells), mask 1'] = (ps, (ps > 0.5).to(torch.uint8), (ps > 0.4).to(torch.uint8), torch.ones(n, dtype=torch.uint8, device=dev))
blk = (torch.arange(n, device=dev) // 7200) % 4 == 0
cases['iid p inside a blocky 25% mask, zeros outside'] = (torch.where(blk, p, torch.zeros_like(p)), ((p > 0.5) & blk).to(torch.uint8),
                                                         cases['iid uniform p, Bernoulli target, 25% mask'][2] * blk.to(torch.uint8), blk.to(torch.uint8))
for name, (pp, pred, target, mask) in cases.items():
    us = timeit(lambda: metrics.eval_fused(pp, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=1, sync=False, break_table=bt))
    print('%-52s %6.1f us  (%.2f of HBM)' % (name, us, 7.0 * n / us / 1e3 / 6449.1), flush=True)
