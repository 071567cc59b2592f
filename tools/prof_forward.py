"""Short driver for ncu captures: a few chunks of the BraTS MC-dropout forward (+ aggregation + fused metrics).

    python tools/prof_forward.py [n_slices=16] [T=20]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import model, steps, metrics, synth  # noqa: E402

torch.set_grad_enabled(False)
n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sd = synth.random_unet_state_dict(in_channels=4, seed=20)
net = model.B200UNet(sd, in_channels=4, dropout=0.05, device='cuda:0', seed=20)
x = torch.randn((n_slices, 4, 240, 240), generator=torch.Generator().manual_seed(1)).cuda()
logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True)
out = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True)
n = out['foreground'].numel()
target = (torch.rand(n, device='cuda') < out['foreground'].view(-1)).to(torch.uint8)
mask = (torch.rand(n, device='cuda') < 0.5).to(torch.uint8)
metrics.eval_fused(out['foreground'], out['prediction'], target, mask, sync=False)
torch.cuda.synchronize()
print('launches', net.last_launch_count())
