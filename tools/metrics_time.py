"""Device time of the fused ECE + U-E histogram pass (1 subject and 50 subjects per launch) and of the aggregation."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import metrics, steps, tables  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
vps = 155 * 240 * 240
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for S in (1, 50):
    n = S * vps
    g = torch.Generator(device=dev).manual_seed(1)
    p = torch.rand(n, device=dev, generator=g)
    target = (torch.rand(n, device=dev, generator=g) < p).to(torch.uint8)
    pred = (p > 0.5).to(torch.uint8)
    mask = (torch.rand(n, device=dev, generator=g) < 0.5).to(torch.uint8)
    ms = timeit(lambda: metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, sync=False, break_table=bt))
    print('eval_fused S=%2d: %.4f ms  (%.1f us/subject, %.0f GB/s of 7 B/voxel)' % (S, ms, ms * 1e3 / S, 7.0 * n / ms / 1e6))
    del p, target, pred, mask
T = 20
logits = torch.randn((T + 1, 155, 240, 240, 2), device=dev)
ms = timeit(lambda: steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True))
print('aggregate T=20 one subject: %.4f ms (%.0f GB/s of %d B/voxel)' % (ms, (8 * T + 12 + 4 + 1 + 4) * vps / ms / 1e6, 8 * T + 21))
ms = timeit(lambda: steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True, ws_logits=logits[0]))
print('aggregate T=20 + weight-scaling softmax in the same launch: %.4f ms (%.0f GB/s of %d B/voxel; %.0f GB/s on the 8T+12 = 172 B/voxel accounting)'
      % (ms, (8 * T + 37) * vps / ms / 1e6, 8 * T + 37, (8 * T + 12) * vps / ms / 1e6))
diff = (logits[..., 0] - logits[..., 1]).contiguous()
ms = timeit(lambda: steps.summarize(steps.LazyMultiProbabilities(diff[1:], diff=True), emit_prediction=True, emit_foreground=True, ws_logits=diff[0]))
print('aggregate T=20 + weight-scaling softmax on logit DIFFERENCES (what the head writes for McPredictStep): %.4f ms (%.0f GB/s of the %d B/voxel '
      'it moves; %.0f GB/s on the 8T+12 = 172 B/voxel accounting)' % (ms, (4 * T + 29) * vps / ms / 1e6, 4 * T + 29, (8 * T + 12) * vps / ms / 1e6))
del diff, logits
for S in (1, 50):   # Beta(0.3, 0.3) maps, target ~ Bernoulli(p), 25 % mask: the data-dependent table reads see a spread
    n = S * vps
    bd = torch.distributions.Beta(torch.tensor(0.3, device=dev), torch.tensor(0.3, device=dev))
    p = bd.sample((n,)).float().clamp_(0, 1)
    target = (torch.rand(n, device=dev) < p).to(torch.uint8)
    pred = (p > 0.5).to(torch.uint8)
    mask = (torch.rand(n, device=dev) < 0.25).to(torch.uint8)
    ms = timeit(lambda: metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, sync=False, break_table=bt))
    print('eval_fused Beta(0.3,0.3) S=%2d: %.4f ms  (%.1f us/subject, %.0f GB/s of 7 B/voxel)' % (S, ms, ms * 1e3 / S, 7.0 * n / ms / 1e6))
    del p, target, pred, mask
