"""ncu driver for the two HBM-side kernels at the BASELINE size: the summary on logit differences of one 155 x 240 x 240 subject
(T = 20 + the weight-scaling sample) and the fused ECE / U-E histogram pass for one and for eight such subjects (Beta(0.3, 0.3)
maps, Bernoulli(p) target, 25 % mask).    ncu --set full -k regex:'aggregate_kernel|eval_fused' python tools/prof_hbm.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import steps, metrics, tables  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
T, vps = 20, 155 * 240 * 240
diff = torch.randn((T + 1, 155, 240, 240), device=dev) * 3
for _ in range(2):
    steps.summarize(steps.LazyMultiProbabilities(diff[1:], diff=True), emit_prediction=True, emit_foreground=True, ws_logits=diff[0])
del diff
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
for S in (1, 8):
    n = S * vps
    g = torch.Generator(device=dev).manual_seed(20)
    u = torch.rand(n, device=dev, generator=g).pow_(1.0 / 0.3)
    v = torch.rand(n, device=dev, generator=g).pow_(1.0 / 0.3)
    p = (u / (u + v + 1e-30)).clamp_(0, 1)
    del u, v
    target = (torch.rand(n, device=dev, generator=g) < p).to(torch.uint8)
    pred = (p > 0.5).to(torch.uint8)
    mask = (torch.rand(n, device=dev, generator=g) < 0.25).to(torch.uint8)
    for _ in range(2):
        metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, sync=False, break_table=bt)
    del p, target, pred, mask
torch.cuda.synchronize()
