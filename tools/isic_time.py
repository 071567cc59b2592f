"""ISIC-shaped throughput (BASELINE.json: 3 x 256 x 256 images): MC dropout T = 20 (+ the weight-scaling pass) forward and
the fused summary over a batch of images resident in HBM.   python tools/isic_time.py [n_images=64]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import model, steps, synth  # noqa: E402

torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 20
net = model.B200UNet(synth.random_unet_state_dict(in_channels=3, seed=20), in_channels=3, dropout=0.05, device='cuda:0', seed=20)
x = torch.rand((n, 3, 256, 256), generator=torch.Generator().manual_seed(1)).cuda()


def step():
    logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True)
    return steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
vox = n * 256 * 256
print('ISIC-shaped MC T=%d: %d images 3x256x256 in %.2f ms -> %.3e voxel-samples/s (%.1f images/s; %.0f TFLOP/s algorithmic at 517952 FLOP/voxel-sample incl. the weight-scaling pass)'
      % (T, n, ms, vox * T / (ms * 1e-3), n / (ms * 1e-3), vox * (T + 1) * 517952 / (ms * 1e-3) / 1e12))
