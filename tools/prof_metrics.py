import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa
from rcu_b200 import metrics, tables
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
vps = 155 * 240 * 240
n = S * vps
dev = 'cuda:0'
p = torch.rand(n, device=dev)
target = (torch.rand(n, device=dev) < p).to(torch.uint8)
pred = (p > 0.5).to(torch.uint8)
mask = (torch.rand(n, device=dev) < 0.25).to(torch.uint8)
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
for _ in range(2):
    metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, sync=False, break_table=bt)
torch.cuda.synchronize()
